#!/bin/bash
# sweep of the compressed share of the host-to-device transfer (e2e leg)
mkdir -p gpurun_out
for share in adaptive 0.45 0.6 0.75 0.9; do
  if [ $share = adaptive ]; then unset BENCH_E2E_PACK; else export BENCH_E2E_PACK=$share; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2e2e_$share.json 2> gpurun_out/r2e2e_$share.err
  python -c "
import json
d = json.loads(open('gpurun_out/r2e2e_$share.json').read().strip().split('\n')[-1]); e = d['e2e']
print('$share: e2e %.1f M reads/s, raw %.1f, ragged %.1f, h2d B/read %.1f, pack_fraction %s, pack_s %.3f total_s %.3f lane_wait %.3f' % (e['value'] / 1e6, e['raw_transfer_value'] / 1e6, e['ragged']['value'] / 1e6, e['h2d_bytes_per_step'] / d['config']['reads_per_step_per_gpu'], e['host_profile']['pack_fraction'], e['host_profile']['pack_s'], e['host_profile']['total_s'], e['host_profile']['lane_wait_s']))"
done
# config 5: where does the host side of the e2e leg spend its time?
export CUTADAPT_B200_HOST_TRACE=1
unset BENCH_E2E_PACK
timeout 600 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2e2e_c5.json 2> gpurun_out/r2e2e_c5.err
grep "process_batch" gpurun_out/r2e2e_c5.err | tail -14 | cut -c1-330
python -c "
import json
d = json.loads(open('gpurun_out/r2e2e_c5.json').read().strip().split('\n')[-1]); e = d['e2e']
print('c5: e2e %.1f M reads/s, raw %.1f, ragged %.1f' % (e['value'] / 1e6, e['raw_transfer_value'] / 1e6, e['ragged']['value'] / 1e6), e['host_profile'])"
