#!/bin/bash
# tools/ab.sh OUT lib1[@ENV=VALUE] lib2 ...   ("default" = the in-tree library): A/B of library variants on the GPU box
out=$1; shift
: > $out
for lib in "$@"; do
  echo "== $lib" >> $out
  envs=${lib#*@}; [ "$envs" = "$lib" ] && envs=""
  lib=${lib%%@*}
  unset CUTADAPT_B200_SUB_READS
  [ -n "$envs" ] && export $envs
  if [ "$lib" = default ]; then unset CUTADAPT_B200_LIB; else export CUTADAPT_B200_LIB=$PWD/build_variants/lib_$lib.so; fi
  timeout 300 python tools/measure_configs.py 4000000 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-60s %8.1f M reads/s' % (d['case'][:60], d['reads_per_s'] / 1e6))" >> $out
  timeout 300 python bench.py --reads ${AB_READS:-20000000} --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1]); print('  bench: value %.1f M reads/s, ms/step %.3f, roofline frac %.4f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac']))" >> $out
done
cat $out
