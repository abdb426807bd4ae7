// cg_host_algos.cpp -- host-side neighbourhood generators used to build the anchored-adapter
// index (adapters.py:1416-1442).  They run once per adapter set, never per read.
//
//   cg_edit_environment     == edit_environment(t, k)       (_align.pyx:785-882)
//   cg_hamming_environment  == hamming_environment(s, k)    (align.py:63-75 over
//                              hamming_sphere, _align.pyx:717-782)
//
// Both enumerate strings over ACGT; each distinct string is reported exactly once together
// with its distance and the number of matching positions of the alignment the reference's
// tie-breaking picks (diagonal first, then left, then up; _align.pyx:838-846).  The order of
// the records is not part of the contract (the index is a dict).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/cutadapt_b200.h"

namespace {

struct Sink {
    int32_t stride;
    uint8_t *strings;
    int32_t *lengths, *errors, *matches;
    int64_t capacity, count;
    void put(const char *s, int len, int e, int m)
    {
        if (count < capacity) {
            if (strings) {
                uint8_t *dst = strings + (size_t)count * stride;
                memset(dst, 0, (size_t)stride);
                memcpy(dst, s, (size_t)(len < stride ? len : stride));
            }
            if (lengths) lengths[count] = len;
            if (errors) errors[count] = e;
            if (matches) matches[count] = m;
        }
        ++count;
    }
};

const char ALPHABET[4] = {'A', 'C', 'G', 'T'};

// Depth-first walk over all strings s; row d of the banded DP aligns s[:d] against t.
struct EditWalker {
    const uint8_t *t; int n, k;
    std::vector<int> cost, match;   // (n+k+1) rows x (n+1) columns
    std::vector<char> s;
    Sink *sink;
    int width() const { return n + 1; }
    static int code(uint8_t c)
    {
        switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1;
                     case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 100 + c; }
    }
    void walk(int depth)
    {
        const int w = width();
        const int *crow = &cost[(size_t)depth * w];
        if (crow[n] <= k) sink->put(s.data(), depth, crow[n], match[(size_t)depth * w + n]);
        if (depth == n + k) return;
        // rows can only get worse: stop when the whole band is already above k (_align.pyx:863-865)
        // (the minimum is taken over the band cells j >= 1 only, like the reference's min_cost)
        int lo = depth - k > 1 ? depth - k : 1, hi = depth + k < n ? depth + k : n, best = depth == 0 ? 0 : 1 << 30;
        for (int j = lo; j <= hi && depth > 0; ++j) best = crow[j] < best ? crow[j] : best;
        if (best > k) return;
        const int i = depth + 1;
        int *nc = &cost[(size_t)i * w], *nm = &match[(size_t)i * w];
        const int *pm = &match[(size_t)depth * w];
        for (int a = 0; a < 4; ++a) {
            const int INF = 1 << 28;
            for (int j = 0; j <= n; ++j) { nc[j] = INF; nm[j] = 0; }
            nc[0] = i;
            const int jlo = i - k > 1 ? i - k : 1, jhi = i + k < n ? i + k : n;
            for (int j = jlo; j <= jhi; ++j) {
                const int mis = code(t[j - 1]) == a ? 0 : 1;
                const int diag = crow[j - 1] + mis, left = nc[j - 1] + 1, up = crow[j] + 1;
                if (diag <= left && diag <= up) { nc[j] = diag; nm[j] = pm[j - 1] + (1 - mis); }
                else if (left <= up) { nc[j] = left; nm[j] = nm[j - 1]; }
                else { nc[j] = up; nm[j] = pm[j]; }
            }
            s[depth] = ALPHABET[a];
            walk(i);
        }
    }
};

void hamming_walk(std::vector<char> &cur, const uint8_t *s, int n, int pos, int left, int used, Sink &sink)
{
    if (pos == n) { sink.put(cur.data(), n, used, n - used); return; }
    // keep s[pos]
    cur[pos] = (char)s[pos];
    hamming_walk(cur, s, n, pos + 1, left, used, sink);
    if (left > 0) {
        for (int a = 0; a < 4; ++a) {
            if (ALPHABET[a] == (char)s[pos]) continue;
            cur[pos] = ALPHABET[a];
            hamming_walk(cur, s, n, pos + 1, left - 1, used + 1, sink);
        }
    }
}

}  // namespace

extern "C" int64_t cg_edit_environment(const uint8_t *t, int32_t n, int32_t k, int32_t stride, uint8_t *strings,
                                       int32_t *lengths, int32_t *errors, int32_t *matches, int64_t capacity)
{
    if (n < 0 || k < 0 || (n > 0 && !t) || stride < 0) return CG_EINVAL;
    for (int i = 0; i < n; ++i) if (t[i] & 0x80) return CG_ENONASCII;
    Sink sink{stride, strings, lengths, errors, matches, capacity < 0 ? 0 : capacity, 0};
    EditWalker w;
    w.t = t; w.n = n; w.k = k; w.sink = &sink;
    w.cost.assign((size_t)(n + k + 1) * (n + 1), 1 << 28);
    w.match.assign((size_t)(n + k + 1) * (n + 1), 0);
    w.s.assign((size_t)n + k + 1, 0);
    for (int j = 0; j <= n; ++j) w.cost[j] = j;   // row 0 (_align.pyx:813-814)
    w.walk(0);
    return sink.count;
}

extern "C" int64_t cg_hamming_environment(const uint8_t *s, int32_t n, int32_t k, int32_t stride, uint8_t *strings,
                                          int32_t *errors, int32_t *matches, int64_t capacity)
{
    if (n < 0 || k < 0 || (n > 0 && !s) || stride < 0) return CG_EINVAL;
    for (int i = 0; i < n; ++i) if (s[i] & 0x80) return CG_ENONASCII;
    Sink sink{stride, strings, nullptr, errors, matches, capacity < 0 ? 0 : capacity, 0};
    std::vector<char> cur((size_t)n + 1, 0);
    hamming_walk(cur, s, n, 0, k, 0, sink);
    return sink.count;
}
