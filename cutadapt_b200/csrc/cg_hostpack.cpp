// cg_hostpack.cpp -- host worker pool and the base-6 read packer (see cg_hostpack.h)
#include "cg_hostpack.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define CG_CPU_RELAX() _mm_pause()
#else
#define CG_CPU_RELAX() ((void)0)
#endif

// CPUs this process may actually use: the affinity mask, cut down by a cgroup CPU quota if there is one
// (a container on a shared host usually sees every hardware thread but may only run on a few).
int cg_host_cpus()
{
    int n = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) { const int c = CPU_COUNT(&set); if (c > 0) n = c; }
    double quota = -1.0;
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {                       // cgroup v2: "<quota|max> <period>"
        char q[64]; long long period = 0;
        if (fscanf(f, "%63s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) quota = atof(q) / (double)period;
        fclose(f);
    } else {
        long long q = -1, period = 0;                                            // cgroup v1
        if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (fscanf(g, "%lld", &q) != 1) q = -1; fclose(g); }
        if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(g, "%lld", &period) != 1) period = 0; fclose(g); }
        if (q > 0 && period > 0) quota = (double)q / (double)period;
    }
    if (quota > 0.0 && quota < n) n = (int)(quota + 0.5);
#endif
    return n < 1 ? 1 : n;
}

int cg_host_threads_default()
{
    if (const char *e = getenv("CUTADAPT_B200_HOST_THREADS")) {
        int v = atoi(e);
        if (v >= 1) return v > 256 ? 256 : v;
    }
    // leave the machine usable: at most half of the usable CPUs, 32 at most; all of a small quota
    // (measured on a 16-CPU quota: 16 workers pack 69 GB/s, 14 workers 49 GB/s).
    const int cpus = cg_host_cpus();
    int n = cpus <= 16 ? cpus : cpus / 2;
    if (n > 32) n = 32;
    return n < 1 ? 1 : n;
}

CgHostPool::CgHostPool(int n_threads)
{
    for (int i = 1; i < n_threads; ++i) workers_.emplace_back([this, i] { worker_main(i); });
}

CgHostPool::~CgHostPool()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_.store(true, std::memory_order_release);
        generation_.fetch_add(1, std::memory_order_release);
    }
    cv_start_.notify_all();
    for (auto &t : workers_) t.join();
}

void CgHostPool::worker_main(int id)
{
    uint64_t seen = 0;
    for (;;) {
        // wait for the next job set: spin first (about a millisecond), then sleep
        int spins = 0;
        while (generation_.load(std::memory_order_acquire) == seen) {
            if (++spins < 20000) { CG_CPU_RELAX(); continue; }
            std::unique_lock<std::mutex> lk(mu_);
            sleepers_.fetch_add(1);   // seq_cst: pairs with run()'s generation_++ / sleepers_ read
            cv_start_.wait(lk, [&] { return generation_.load() != seen; });
            sleepers_.fetch_sub(1, std::memory_order_acq_rel);
        }
        seen = generation_.load(std::memory_order_acquire);
        if (stop_.load(std::memory_order_acquire)) return;
        const std::function<void(int64_t, int)> *fn = fn_;
        const int64_t n_jobs = n_jobs_.load(std::memory_order_acquire);
        for (;;) {
            const int64_t j = next_.fetch_add(1, std::memory_order_relaxed);
            if (j >= n_jobs) break;
            (*fn)(j, id);
        }
        pending_.fetch_sub(1, std::memory_order_acq_rel);
    }
}

void CgHostPool::run(int64_t n_jobs, const std::function<void(int64_t, int)> &fn)
{
    if (n_jobs <= 0) return;
    if (workers_.empty() || n_jobs == 1) {
        for (int64_t j = 0; j < n_jobs; ++j) fn(j, 0);
        return;
    }
    fn_ = &fn;
    n_jobs_.store(n_jobs, std::memory_order_relaxed);
    next_.store(0, std::memory_order_relaxed);
    pending_.store((int)workers_.size(), std::memory_order_relaxed);
    generation_.fetch_add(1);                                  // publishes the job set (seq_cst)
    if (sleepers_.load() > 0) {
        { std::lock_guard<std::mutex> lk(mu_); }
        cv_start_.notify_all();
    }
    for (;;) {
        const int64_t j = next_.fetch_add(1, std::memory_order_relaxed);
        if (j >= n_jobs) break;
        fn(j, 0);
    }
    // every worker checks in (it may still be inside its last job)
    int spins = 0;
    while (pending_.load(std::memory_order_acquire) != 0) {
        if (++spins < 4000) CG_CPU_RELAX();
        else std::this_thread::yield();
    }
    fn_ = nullptr;
}

int CgHostPool::follow_memory(const void *addr)
{
    if (followed_node_ != -2) return followed_node_;
    followed_node_ = -1;
#if defined(__linux__) && defined(__x86_64__)
    if (const char *e = getenv("CUTADAPT_B200_NUMA")) if (e[0] == '0') return -1;
    if (workers_.empty()) return -1;
    // which node holds the page?  get_mempolicy(&node, NULL, 0, addr, MPOL_F_NODE | MPOL_F_ADDR)
    int node = -1;
    if (syscall(SYS_get_mempolicy, &node, nullptr, 0UL, addr, 3UL) != 0 || node < 0) return -1;
    char path[96];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char list[1024] = {0};
    const bool ok = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!ok) return -1;
    // is there more than one node at all?
    if (FILE *g = fopen("/sys/devices/system/node/node1/cpulist", "r")) fclose(g); else return -1;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n_cpus = 0;
    for (char *p = list; *p;) {                       // "0-31,64-95"
        char *end;
        const long a = strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        if (*end == '-') { p = end + 1; b = strtol(p, &end, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, &set); ++n_cpus; }
        p = *end == ',' ? end + 1 : end;
        if (*end != ',') break;
    }
    if (n_cpus < (int)workers_.size()) return -1;     // do not squeeze the pool onto fewer CPUs than workers
    for (auto &t : workers_) pthread_setaffinity_np(t.native_handle(), sizeof set, &set);
    followed_node_ = node;
#else
    (void)addr;
#endif
    return followed_node_;
}

// ------------------------------------------------------------------------------------------
// base-6 packer
// ------------------------------------------------------------------------------------------
namespace {
struct PackTables {
    uint16_t t0[256], t1[256], t2[256];   // code * 36 / * 6 / * 1, bit 8.. = escape marker
    uint8_t cls[256];
    PackTables()
    {
        for (int c = 0; c < 256; ++c) {
            int v = CG_PACK_ESCAPE;
            switch (c) {
            case 'A': v = 0; break;
            case 'C': v = 1; break;
            case 'G': v = 2; break;
            case 'T': v = 3; break;
            case 'N': v = 4; break;
            default: break;
            }
            cls[c] = (uint8_t)v;
            const uint16_t esc = v == CG_PACK_ESCAPE ? 0x100 : 0;
            t0[c] = (uint16_t)(v * 36) | esc;
            t1[c] = (uint16_t)(v * 6) | esc;
            t2[c] = (uint16_t)v | esc;
        }
    }
};
const PackTables g_pack;
}  // namespace

#if defined(__x86_64__)
// AVX2 body of the packer: 24 characters -> 8 stream bytes per step.  The class of a character is
// looked up by its low nibble (A=1, C=3, G=7, T=4, N=E are distinct) and confirmed by comparing
// with the expected character; everything else becomes the escape code.  Returns the number of
// stream bytes written (a multiple of 8); *any_escape is set if an escape code was produced.
__attribute__((target("avx2"))) static int64_t pack3_avx2(const uint8_t *s, int64_t n_out, uint8_t *dst,
                                                           bool *any_escape)
{
    const __m256i nib = _mm256_set1_epi8(0x0F);
    const __m256i lut_code = _mm256_setr_epi8(5, 0, 5, 1, 3, 5, 5, 2, 5, 5, 5, 5, 5, 5, 4, 5,
                                              5, 0, 5, 1, 3, 5, 5, 2, 5, 5, 5, 5, 5, 5, 4, 5);
    const __m256i lut_char = _mm256_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, 'N', -1,
                                              -1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, 'N', -1);
    const __m256i five = _mm256_set1_epi8(CG_PACK_ESCAPE);
    const __m256i spread = _mm256_setr_epi8(0, 1, 2, -128, 3, 4, 5, -128, 6, 7, 8, -128, 9, 10, 11, -128,
                                            0, 1, 2, -128, 3, 4, 5, -128, 6, 7, 8, -128, 9, 10, 11, -128);
    const __m256i weights = _mm256_setr_epi8(36, 6, 1, 0, 36, 6, 1, 0, 36, 6, 1, 0, 36, 6, 1, 0,
                                             36, 6, 1, 0, 36, 6, 1, 0, 36, 6, 1, 0, 36, 6, 1, 0);
    const __m256i ones = _mm256_set1_epi16(1);
    __m256i esc = _mm256_setzero_si256();
    int64_t k = 0;
    for (; k + 8 <= n_out; k += 8, s += 24) {
        const __m128i a = _mm_loadu_si128((const __m128i *)s);
        const __m128i b = _mm_loadu_si128((const __m128i *)(s + 12));
        const __m256i x = _mm256_inserti128_si256(_mm256_castsi128_si256(a), b, 1);
        const __m256i lo = _mm256_and_si256(x, nib);
        const __m256i cand = _mm256_shuffle_epi8(lut_code, lo);
        const __m256i expect = _mm256_shuffle_epi8(lut_char, lo);
        const __m256i ok = _mm256_cmpeq_epi8(x, expect);
        const __m256i code = _mm256_blendv_epi8(five, cand, ok);
        const __m256i t = _mm256_shuffle_epi8(code, spread);          // c0 c1 c2 0 | c3 c4 c5 0 | ...
        esc = _mm256_or_si256(esc, _mm256_cmpeq_epi8(t, five));
        const __m256i m1 = _mm256_maddubs_epi16(t, weights);           // 36 c0 + 6 c1, c2
        const __m256i m2 = _mm256_madd_epi16(m1, ones);                // 32-bit: stream byte value
        const __m256i p16 = _mm256_packus_epi32(m2, m2);
        const __m256i p8 = _mm256_packus_epi16(p16, p16);
        const uint32_t w0 = (uint32_t)_mm256_cvtsi256_si32(p8);
        const uint32_t w1 = (uint32_t)_mm256_extract_epi32(p8, 4);
        memcpy(dst + k, &w0, 4);
        memcpy(dst + k + 4, &w1, 4);
    }
    *any_escape = !_mm256_testz_si256(esc, esc);
    return k;
}
struct alignas(64) Pack512Tables {
    uint8_t idx[64], w[64], code[64], chr[64];
    Pack512Tables()
    {
        static const uint8_t code16[16] = {5, 0, 5, 1, 3, 5, 5, 2, 5, 5, 5, 5, 5, 5, 4, 5};
        static const uint8_t char16[16] = {255, 'A', 255, 'C', 'T', 255, 255, 'G', 255, 255, 255, 255, 255, 255, 'N', 255};
        for (int t = 0; t < 16; ++t) {
            for (int b = 0; b < 4; ++b) idx[4 * t + b] = (uint8_t)(b < 3 ? 3 * t + b : 0);
            w[4 * t] = 36; w[4 * t + 1] = 6; w[4 * t + 2] = 1; w[4 * t + 3] = 0;
        }
        for (int i = 0; i < 64; ++i) { code[i] = code16[i & 15]; chr[i] = char16[i & 15]; }
    }
};
// AVX-512 (BW + VBMI) body: 48 characters -> 16 stream bytes per step, same scheme; one byte permutation
// spreads the triples to 4-byte groups (c0 c1 c2 0).  Loads 64 bytes per step.
__attribute__((target("avx512f,avx512bw,avx512vbmi"))) static int64_t pack3_avx512(const uint8_t *s, int64_t n_out,
                                                                                    uint8_t *dst, bool *any_escape)
{
    static const Pack512Tables T;
    const uint8_t *idx_b = T.idx, *w_b = T.w, *code_b = T.code, *char_b = T.chr;
    const __m512i idx = _mm512_load_si512(idx_b), weights = _mm512_load_si512(w_b);
    const __m512i lut_code = _mm512_load_si512(code_b), lut_char = _mm512_load_si512(char_b);
    const __m512i nib = _mm512_set1_epi8(0x0F), five = _mm512_set1_epi8(CG_PACK_ESCAPE);
    const __m512i ones = _mm512_set1_epi16(1);
    const __mmask64 valid = 0x7777777777777777ull;
    __mmask64 esc = 0;
    int64_t k = 0;
    for (; k + 16 <= n_out; k += 16, s += 48) {
        _mm_prefetch((const char *)s + 1536, _MM_HINT_T0);   // the hardware prefetcher stops at page ends
        const __m512i x = _mm512_permutexvar_epi8(idx, _mm512_loadu_si512((const void *)s));
        const __m512i lo = _mm512_and_si512(x, nib);
        const __m512i cand = _mm512_shuffle_epi8(lut_code, lo);
        const __m512i expect = _mm512_shuffle_epi8(lut_char, lo);
        const __mmask64 ok = _mm512_cmpeq_epi8_mask(x, expect);
        const __m512i code = _mm512_mask_blend_epi8(ok, five, cand);     // weight 0 hides the filler bytes
        esc |= ~ok & valid;
        const __m512i m1 = _mm512_maddubs_epi16(code, weights);
        const __m512i m2 = _mm512_madd_epi16(m1, ones);
        _mm_storeu_si128((__m128i *)(dst + k), _mm512_cvtepi32_epi8(m2));
    }
    *any_escape = esc != 0;
    return k;
}
static const bool g_have_avx2 = __builtin_cpu_supports("avx2");
static const bool g_have_avx512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
                                  !getenv("CUTADAPT_B200_NO_AVX512");
#else
static const bool g_have_avx2 = false, g_have_avx512 = false;
static int64_t pack3_avx2(const uint8_t *, int64_t, uint8_t *, bool *) { return 0; }
static int64_t pack3_avx512(const uint8_t *, int64_t, uint8_t *, bool *) { return 0; }
#endif

void cg_pack3_range(const uint8_t *seq, int64_t a0, int64_t lo, int64_t hi, int64_t i0, int64_t i1,
                    uint8_t *packed, std::vector<uint64_t> &exc)
{
    const PackTables &T = g_pack;
    int64_t i = i0;
    // triples that touch positions outside [lo, hi): byte-wise with bounds checks
    auto slow = [&](int64_t k) {
        unsigned v = 0;
        static const int mul[3] = {36, 6, 1};
        for (int b = 0; b < 3; ++b) {
            const int64_t pos = a0 + 3 * k + b;
            if (pos < lo || pos >= hi) continue;   // filler 'A' = 0
            const uint8_t c = seq[pos];
            const unsigned code = T.cls[c];
            v += code * mul[b];
            if (code == CG_PACK_ESCAPE) exc.push_back(((uint64_t)(pos - a0) << 8) | c);
        }
        packed[k] = (uint8_t)v;
    };
    // first triple fully inside: a0 + 3k >= lo  ->  k >= ceil((lo - a0) / 3)
    int64_t k_in0 = lo <= a0 ? 0 : (lo - a0 + 2) / 3;
    // triples fully inside end before: a0 + 3k + 2 < hi  ->  k < floor((hi - a0) / 3)
    int64_t k_in1 = hi - a0 >= 3 ? (hi - a0) / 3 : 0;
    if (k_in0 < i0) k_in0 = i0;
    if (k_in0 > i1) k_in0 = i1;
    if (k_in1 > i1) k_in1 = i1;
    if (k_in1 < k_in0) k_in1 = k_in0;
    for (; i < k_in0; ++i) slow(i);
    // fast part in blocks: escapes are detected per block and resolved by a second look
    const int64_t BLOCK = 4096;
    while (i < k_in1) {
        const int64_t e = i + BLOCK < k_in1 ? i + BLOCK : k_in1;
        const uint8_t *s = seq + a0 + 3 * i;
        unsigned any = 0;
        int64_t k = i;
        if (g_have_avx2) {
            // the vector bodies read 28 / 64 bytes per step of 24 / 48: keep them 4 / 16 bytes away from `hi`
            const int64_t k_safe = (hi - a0 - (g_have_avx512 ? 16 : 4)) / 3;
            const int64_t e_v = e < k_safe ? e : k_safe;
            if (e_v - i >= 16) {
                bool escaped = false;
                const int64_t done = g_have_avx512 ? pack3_avx512(s, e_v - i, packed + i, &escaped)
                                                   : pack3_avx2(s, e_v - i, packed + i, &escaped);
                if (escaped) any |= 0x100;
                k += done;
                s += 3 * done;
            }
        }
        for (; k < e; ++k, s += 3) {
            const unsigned v = (unsigned)T.t0[s[0]] + T.t1[s[1]] + T.t2[s[2]];
            packed[k] = (uint8_t)v;
            any |= v;
        }
        if (any >> 8) {
            const uint8_t *b = seq + a0 + 3 * i, *be = seq + a0 + 3 * e;
            for (; b < be; ++b)
                if (T.cls[*b] == CG_PACK_ESCAPE) exc.push_back(((uint64_t)(b - (seq + a0)) << 8) | *b);
        }
        i = e;
    }
    for (; i < i1; ++i) slow(i);
}
