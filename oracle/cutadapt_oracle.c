/*
 * oracle/cutadapt_oracle.c -- CPU restatement of the cutadapt adapter-trimming hot path.
 *
 * TEST INFRASTRUCTURE.  This file is the *checker*, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The shipped library (cutadapt_b200/libcutadapt_b200.so) does not link,
 * include or call anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 *   (a) the committed golden vectors in tests/golden/ (generated in the build container
 *       from the reference's own compiled Cython, script tests/golden/make_golden.py), and
 *   (b) where oracle/_ref/ is present, the reference itself on randomized inputs.
 *
 * All file:line citations are relative to the reference checkout (marcelm/cutadapt).
 * Plain C99, no dependencies:  gcc -O2 -shared -fPIC -o liboracle.so cutadapt_oracle.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* Character classes (src/cutadapt/_match_tables.py:4-66)                                */
/* ------------------------------------------------------------------------------------ */

enum { ENC_UPPER = 0, ENC_ACGT = 1, ENC_IUPAC = 2 };

static unsigned char tab_upper[256], tab_acgt[256], tab_iupac[256];
static int tables_ready = 0;

static void put_both_cases(unsigned char *t, char c, unsigned char v)
{
    t[(unsigned char)c] = v;
    t[(unsigned char)(c | 0x20)] = v;
}

static void init_tables(void)
{
    int c;
    if (tables_ready) return;
    /* _upper_table(): bytes(range(256)).upper()  -- only a..z change (_match_tables.py:64-66) */
    for (c = 0; c < 256; c++) tab_upper[c] = (unsigned char)((c >= 'a' && c <= 'z') ? c - 32 : c);
    /* _acgt_table(): A=1 C=2 G=4 T/U=8, everything else 0x80 (_match_tables.py:4-17) */
    memset(tab_acgt, 0x80, 256);
    put_both_cases(tab_acgt, 'A', 1); put_both_cases(tab_acgt, 'C', 2);
    put_both_cases(tab_acgt, 'G', 4); put_both_cases(tab_acgt, 'T', 8);
    put_both_cases(tab_acgt, 'U', 8);
    /* _iupac_table(): bit sets; N additionally carries 0x80; X = 0 (_match_tables.py:20-61) */
    memset(tab_iupac, 0, 256);
    {
        static const struct { char c; unsigned char v; } iu[] = {
            {'X', 0}, {'A', 1}, {'C', 2}, {'G', 4}, {'T', 8}, {'U', 8},
            {'R', 1 | 4}, {'Y', 2 | 8}, {'S', 4 | 2}, {'W', 1 | 8}, {'K', 4 | 8}, {'M', 1 | 2},
            {'B', 2 | 4 | 8}, {'D', 1 | 4 | 8}, {'H', 1 | 2 | 8}, {'V', 1 | 2 | 4},
            {'N', 0x8F},
        };
        size_t q;
        for (q = 0; q < sizeof iu / sizeof iu[0]; q++) put_both_cases(tab_iupac, iu[q].c, iu[q].v);
    }
    tables_ready = 1;
}

static const unsigned char *table_for(int enc)
{
    init_tables();
    return enc == ENC_IUPAC ? tab_iupac : enc == ENC_ACGT ? tab_acgt : tab_upper;
}

/* translate() (_align.pyx:43-56): returns -2 for a non-ASCII byte, like the ValueError there */
static int encode(const unsigned char *s, int n, int enc, unsigned char *out)
{
    const unsigned char *t = table_for(enc);
    int i;
    for (i = 0; i < n; i++) {
        if (s[i] & 0x80) return -2;
        out[i] = t[s[i]];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Aligner.locate (_align.pyx:298-587)                                                   */
/* ------------------------------------------------------------------------------------ */

typedef struct { int cost, score, origin; } cell_t;          /* _align.pyx:23-26 */

/*
 * Returns 1 and fills out[6] = (ref_start, ref_stop, query_start, query_stop, score, errors),
 * 0 for "no match" (Python None), negative on error (-1 bad argument, -2 non-ASCII,
 * -4 out of memory).
 *
 * Reference encoding follows Aligner._set_reference (_align.pyx:250-277); the query encoding
 * and the compare mode follow locate() (_align.pyx:322-329).
 */
int oracle_locate(const unsigned char *ref_ascii, int m,
                  const unsigned char *query_ascii, int n,
                  double max_error_rate, int flags,
                  int wildcard_ref, int wildcard_query,
                  int indel_cost, int min_overlap, int *out)
{
    const int start_in_ref = flags & 1, start_in_query = flags & 2;
    const int stop_in_ref = flags & 4, stop_in_query = flags & 8;
    const int MATCH = 1, MISMATCH = -1, INS_SCORE = -2, DEL_SCORE = -2;  /* _align.pyx:16-19 */
    unsigned char *s1 = NULL, *s2 = NULL;
    cell_t *col = NULL;
    int *n_counts = NULL;
    int effective_length, compare_ascii, i, j, k, min_n, max_n, last, last_filled_i = 0;
    int origin = 0, cost, score, length, cur_eff, best_length, first_i, rc = 0;
    int b_origin, b_cost, b_score, b_ref_stop, b_query_stop;

    if (indel_cost < 1 || m < 0 || n < 0) return -1;            /* _align.pyx:217-218 */
    s1 = malloc((size_t)m + 1); s2 = malloc((size_t)n + 1);
    col = malloc(((size_t)m + 1) * sizeof *col);
    n_counts = malloc(((size_t)m + 1) * sizeof *n_counts);
    if (!s1 || !s2 || !col || !n_counts) { rc = -4; goto done; }

    /* prefix counts of N/n in the reference (_align.pyx:260-267) */
    {
        int c = 0;
        for (i = 0; i < m; i++) {
            n_counts[i] = c;
            if (ref_ascii[i] == 'N' || ref_ascii[i] == 'n') c++;
        }
        n_counts[m] = c;
    }
    effective_length = m;
    if (wildcard_ref) {                                           /* _align.pyx:268-272 */
        effective_length = m - n_counts[m];
        if (effective_length == 0) { rc = -1; goto done; }
        rc = encode(ref_ascii, m, ENC_IUPAC, s1);
    } else if (wildcard_query) {                                  /* _align.pyx:273-274 */
        rc = encode(ref_ascii, m, ENC_ACGT, s1);
    } else {                                                      /* _align.pyx:275-276: raw bytes */
        for (i = 0; i < m; i++) { if (ref_ascii[i] & 0x80) { rc = -2; break; } s1[i] = ref_ascii[i]; }
    }
    if (rc) goto done;
    compare_ascii = 0;
    if (wildcard_query) rc = encode(query_ascii, n, ENC_IUPAC, s2);        /* _align.pyx:322-323 */
    else if (wildcard_ref) rc = encode(query_ascii, n, ENC_ACGT, s2);      /* _align.pyx:324-325 */
    else { rc = encode(query_ascii, n, ENC_UPPER, s2); compare_ascii = 1; }/* _align.pyx:326-328 */
    if (rc) goto done;

    k = (int)(max_error_rate * m);                                /* _align.pyx:343 */
    max_n = n; min_n = 0;                                         /* _align.pyx:346-352 */
    if (!start_in_query) max_n = n < m + k ? n : m + k;
    if (!stop_in_query) min_n = n - m - k > 0 ? n - m - k : 0;

    /* first column, four boundary cases (_align.pyx:364-383) */
    for (i = 0; i <= m; i++) {
        if (!start_in_ref && !start_in_query) {
            col[i].score = i * DEL_SCORE;
            col[i].cost = (i > min_n ? i : min_n) * indel_cost;
            col[i].origin = 0;
        } else if (start_in_ref && !start_in_query) {
            col[i].score = 0;
            col[i].cost = min_n * indel_cost;
            col[i].origin = min_n - i < 0 ? min_n - i : 0;
        } else if (!start_in_ref && start_in_query) {
            col[i].score = i * DEL_SCORE;
            col[i].cost = i * indel_cost;
            col[i].origin = min_n - i > 0 ? min_n - i : 0;
        } else {
            col[i].score = 0;
            col[i].cost = (i < min_n ? i : min_n) * indel_cost;
            col[i].origin = min_n - i;
        }
    }
    /* sentinel "nothing found yet" (_align.pyx:391-396) */
    b_ref_stop = m; b_query_stop = n; b_cost = m + n + 1; b_origin = 0; b_score = 0;

    last = m < k + 1 ? m : k + 1;                                 /* _align.pyx:399-401 */
    if (start_in_ref) last = m;

    for (j = min_n + 1; j <= max_n; j++) {                        /* _align.pyx:433 */
        cell_t diag = col[0];
        /* row 0 (_align.pyx:438-440, increments 413-415) */
        if (start_in_query) col[0].origin += 1;
        else { col[0].cost += indel_cost; col[0].score += INS_SCORE; }
        for (i = 1; i <= last; i++) {                             /* _align.pyx:441-483 */
            int eq = compare_ascii ? (s1[i - 1] == s2[j - 1]) : ((s1[i - 1] & s2[j - 1]) != 0);
            if (eq) {                                             /* _align.pyx:446-453 */
                cost = diag.cost; origin = diag.origin; score = diag.score + MATCH;
            } else {                                              /* _align.pyx:455-476 */
                cell_t cur = col[i], prev = col[i - 1];
                int c_diag = diag.cost + 1;
                int c_ins = cur.cost + indel_cost;
                int c_del = prev.cost + indel_cost;
                if (c_diag <= c_del && c_diag <= c_ins) {
                    cost = c_diag; origin = diag.origin; score = diag.score + MISMATCH;
                } else if (c_del <= c_ins) {
                    cost = c_del; origin = prev.origin; score = prev.score + DEL_SCORE;
                } else {
                    cost = c_ins; origin = cur.origin; score = cur.score + INS_SCORE;
                }
            }
            diag = col[i];
            col[i].cost = cost; col[i].origin = origin; col[i].score = score;
        }
        last_filled_i = last;                                     /* _align.pyx:484 */
        while (last >= 0 && col[last].cost > k) last--;           /* _align.pyx:490-491 */
        if (last < m) {
            last++;                                               /* _align.pyx:494-495 */
        } else if (stop_in_query) {                               /* _align.pyx:496-533 */
            int acceptable;
            cost = col[m].cost; score = col[m].score; origin = col[m].origin;
            length = m + (origin < 0 ? origin : 0);
            cur_eff = length;
            if (wildcard_ref) {
                if (length < m) cur_eff = length - (n_counts[m] - n_counts[m - length]);
                else cur_eff = effective_length;
            }
            acceptable = length >= min_overlap && (double)cost <= cur_eff * max_error_rate;
            best_length = m + (b_origin < 0 ? b_origin : 0);
            if (acceptable && (b_cost == m + n + 1
                               || (origin <= b_origin + m / 2 && score > b_score)
                               || (length > best_length && score > b_score))) {
                b_score = score; b_cost = cost; b_origin = origin;
                b_ref_stop = m; b_query_stop = j;
                if (cost == 0 && origin >= 0) break;              /* _align.pyx:531-533 */
            }
        }
    }

    if (max_n == n) {                                             /* _align.pyx:536-572 */
        first_i = stop_in_ref ? 0 : m;
        for (i = last_filled_i; i >= first_i; i--) {
            int acceptable;
            length = i + (col[i].origin < 0 ? col[i].origin : 0);
            cost = col[i].cost; score = col[i].score;
            if (wildcard_ref) {
                if (length < m) {
                    int ref_start = -(col[i].origin < 0 ? col[i].origin : 0);
                    cur_eff = length - (n_counts[i] - n_counts[ref_start]);
                } else cur_eff = effective_length;
            } else cur_eff = length;
            acceptable = length >= min_overlap && (double)cost <= cur_eff * max_error_rate;
            best_length = b_ref_stop + (b_origin < 0 ? b_origin : 0);
            /* NB: `origin` below is the function-scope variable left over from the column
               loop, not col[i].origin -- exactly as in _align.pyx:565. */
            if (acceptable && (b_cost == m + n + 1
                               || (origin <= b_origin + m / 2 && score > b_score)
                               || (length > best_length && score > b_score))) {
                b_score = score; b_cost = cost; b_origin = col[i].origin;
                b_ref_stop = i; b_query_stop = n;
            }
        }
    }
    if (b_cost == m + n + 1) { rc = 0; goto done; }               /* _align.pyx:573-577 */
    if (b_origin >= 0) { out[0] = 0; out[2] = b_origin; }         /* _align.pyx:579-587 */
    else { out[0] = -b_origin; out[2] = 0; }
    out[1] = b_ref_stop; out[3] = b_query_stop; out[4] = b_score; out[5] = b_cost;
    rc = 1;
done:
    free(s1); free(s2); free(col); free(n_counts);
    return rc;
}

/* Aligner.effective_length (_align.pyx:260-271); -1 if it would be 0 with wildcard_ref */
int oracle_effective_length(const unsigned char *ref_ascii, int m, int wildcard_ref)
{
    int i, c = 0;
    if (!wildcard_ref) return m;
    for (i = 0; i < m; i++) if (ref_ascii[i] == 'N' || ref_ascii[i] == 'n') c++;
    return m - c == 0 ? -1 : m - c;
}

/* ------------------------------------------------------------------------------------ */
/* PrefixComparer / SuffixComparer (_align.pyx:594-714)                                  */
/* ------------------------------------------------------------------------------------ */

static int comparer_setup(const unsigned char *ref, int m, double rate, int wildcard_ref,
                          int wildcard_query, int min_overlap, unsigned char *enc_ref,
                          int *max_k)
{
    int eff = m, i, rc;
    if (wildcard_ref) {                                           /* _align.pyx:627-630 */
        int big = 0, small = 0;
        for (i = 0; i < m; i++) { big += ref[i] == 'N'; small += ref[i] == 'n'; }
        eff -= big - small;                /* sic: count('N') - count('n'), _align.pyx:628 */
        if (eff == 0) return -1;
    }
    if (!(0.0 <= rate && rate <= 1.0)) return -1;                 /* _align.pyx:631-632 */
    *max_k = (int)(rate * eff);                                   /* _align.pyx:633 */
    if (min_overlap < 1) return -1;                               /* _align.pyx:634-635 */
    rc = encode(ref, m, wildcard_ref ? ENC_IUPAC : wildcard_query ? ENC_ACGT : ENC_UPPER, enc_ref);
    return rc;                                                    /* _align.pyx:637-642 */
}

static int compare_core(const unsigned char *r, int m, const unsigned char *q, int n,
                        int compare_ascii, int max_k, int min_overlap, int *length_out,
                        int *errors_out)
{
    int length = m < n ? m : n, errors = 0, i;                    /* _align.pyx:667 */
    for (i = 0; i < length; i++)                                  /* _align.pyx:681-688 */
        errors += compare_ascii ? (r[i] != q[i]) : ((r[i] & q[i]) == 0);
    if (errors > max_k || length < min_overlap) return 0;         /* _align.pyx:690-691 */
    *length_out = length; *errors_out = errors;
    return 1;
}

int oracle_prefix_compare(const unsigned char *ref, int m, const unsigned char *query, int n,
                          double rate, int wildcard_ref, int wildcard_query, int min_overlap,
                          int *out)
{
    unsigned char *r = malloc((size_t)m + 1), *q = malloc((size_t)n + 1);
    int max_k = 0, length = 0, errors = 0, rc;
    if (!r || !q) { free(r); free(q); return -4; }
    rc = comparer_setup(ref, m, rate, wildcard_ref, wildcard_query, min_overlap, r, &max_k);
    if (!rc) rc = encode(query, n, wildcard_query ? ENC_IUPAC : wildcard_ref ? ENC_ACGT : ENC_UPPER, q);
    if (!rc) {
        rc = compare_core(r, m, q, n, !(wildcard_ref || wildcard_query), max_k, min_overlap,
                          &length, &errors);
        if (rc == 1) {                                            /* _align.pyx:692-693 */
            out[0] = 0; out[1] = length; out[2] = 0; out[3] = length;
            out[4] = (length - errors) - errors; out[5] = errors;
        }
    }
    free(r); free(q);
    return rc;
}

int oracle_suffix_compare(const unsigned char *ref, int m, const unsigned char *query, int n,
                          double rate, int wildcard_ref, int wildcard_query, int min_overlap,
                          int *out)
{
    /* SuffixComparer reverses reference and query, then re-bases (_align.pyx:696-714) */
    unsigned char *rr = malloc((size_t)m + 1), *qq = malloc((size_t)n + 1);
    int i, rc, tmp[6];
    if (!rr || !qq) { free(rr); free(qq); return -4; }
    for (i = 0; i < m; i++) rr[i] = ref[m - 1 - i];
    for (i = 0; i < n; i++) qq[i] = query[n - 1 - i];
    rc = oracle_prefix_compare(rr, m, qq, n, rate, wildcard_ref, wildcard_query, min_overlap, tmp);
    if (rc == 1) {
        int length = tmp[1];
        out[0] = m - length; out[1] = m; out[2] = n - length; out[3] = n;
        out[4] = tmp[4]; out[5] = tmp[5];
    }
    free(rr); free(qq);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* KmerFinder (_kmer_finder.pyx:58-63, 106-165, 170-213, 226-257)                        */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int64_t search_start;      /* may be negative: relative to the end */
    int64_t search_stop;       /* 0 = to the end, negative = relative to the end */
    uint64_t init_mask;
    uint64_t found_mask;
} oracle_kmer_entry;           /* mask table entry e lives at masks[128*e .. 128*e+127] */

/* matches_lookup()/populate_needle_mask(): which ASCII codes does pattern char `pc` match
   (_match_tables.py:69-98, _kmer_finder.pyx:220-238).  NUL never matches. */
static int pattern_char_matches(unsigned char pc, int code, int ref_wildcards, int query_wildcards)
{
    init_tables();
    if (code == 0) return 0;
    if (!ref_wildcards && !query_wildcards) return tab_upper[pc] == tab_upper[code];
    if (ref_wildcards && !query_wildcards) return (tab_iupac[pc] & tab_acgt[code]) != 0;
    if (!ref_wildcards && query_wildcards) return (tab_acgt[pc] & tab_iupac[code]) != 0;
    return (tab_iupac[pc] & tab_iupac[code]) != 0;
}

/*
 * Pack one (start, stop, kmers) search set into as many 64-bit words as needed
 * (_kmer_finder.pyx:121-164).  `kmers` is `n_kmers` NUL-terminated strings back to back.
 * Returns the number of entries written (entries/masks must have room for n_kmers of them),
 * or -1 if a k-mer is longer than 64 / not ASCII (the ValueError cases, lines 135-140).
 */
int oracle_kmer_pack(int64_t start, int64_t stop_or_zero, const char *kmers, int n_kmers,
                     int ref_wildcards, int query_wildcards,
                     oracle_kmer_entry *entries, uint64_t *masks)
{
    int produced = 0, index = 0;
    const char *p = kmers;
    while (index < n_kmers) {
        char word[64];
        size_t offset = 0;
        uint64_t init = 0, found = 0, *mk = masks + 128 * (size_t)produced;
        size_t i;
        int code;
        memset(word, 0, sizeof word);
        while (index < n_kmers) {
            size_t len = strlen(p);
            for (i = 0; i < len; i++) if (p[i] & 0x80) return -1;
            if (len > 64) return -1;
            if (offset + len > 64) break;
            /* shift counts modulo 64: what the x86-64 build of the reference does for the
               degenerate empty k-mer (1ULL << (0 + 0 - 1), _kmer_finder.pyx:143-147) */
            init |= 1ULL << (offset & 63);
            memcpy(word + offset, p, len);
            found |= 1ULL << ((offset + len - 1) & 63);
            offset += len;
            p += len + 1;
            index++;
        }
        entries[produced].search_start = start;
        entries[produced].search_stop = stop_or_zero;
        entries[produced].init_mask = init;
        entries[produced].found_mask = found;
        memset(mk, 0, 128 * sizeof *mk);
        for (i = 0; i < offset; i++) {
            if (word[i] == 0) continue;
            for (code = 0; code < 128; code++)
                if (pattern_char_matches((unsigned char)word[i], code, ref_wildcards, query_wildcards))
                    mk[code] |= 1ULL << i;
        }
        produced++;
    }
    return produced;
}

/* kmers_present (_kmer_finder.pyx:170-213) with shift_and_multiple_is_present (241-257).
   Returns 1/0, or -2 for a non-ASCII sequence. */
int oracle_kmers_present(const oracle_kmer_entry *entries, const uint64_t *masks, int n_entries,
                         const unsigned char *seq, int64_t n)
{
    int e;
    int64_t i;
    for (i = 0; i < n; i++) if (seq[i] & 0x80) return -2;
    for (e = 0; e < n_entries; e++) {
        int64_t start = entries[e].search_start, stop = entries[e].search_stop;
        const uint64_t *mk = masks + 128 * (size_t)e;
        uint64_t R = 0;
        if (start < 0) { start += n; if (start < 0) start = 0; }
        else if (start > n) continue;
        if (stop < 0) { stop += n; if (stop <= 0) continue; }
        else if (stop == 0) stop = n;
        if (stop > n) stop = n;   /* the reference reads past the string here (UB); clamp */
        if (stop - start <= 0) continue;
        for (i = start; i < stop; i++) {
            R <<= 1; R |= entries[e].init_mask; R &= mk[seq[i]];
            if (R & entries[e].found_mask) return 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* quality_trim_index (qualtrim.pyx:22-73)                                               */
/* ------------------------------------------------------------------------------------ */

void oracle_quality_trim_index(const unsigned char *qual, int n, int cutoff_front,
                               int cutoff_back, int base, int *start_out, int *stop_out)
{
    int s = 0, best = 0, start = 0, stop = n, i;
    for (i = 0; i < n; i++) {                                     /* qualtrim.pyx:51-59 */
        s += cutoff_front - ((signed char)qual[i] - base);
        if (s < 0) break;
        if (s > best) { best = s; start = i + 1; }
    }
    best = 0; s = 0;
    for (i = n - 1; i >= 0; i--) {                                /* qualtrim.pyx:62-70 */
        s += cutoff_back - ((signed char)qual[i] - base);
        if (s < 0) break;
        if (s > best) { best = s; stop = i; }
    }
    if (start >= stop) { start = 0; stop = 0; }                   /* qualtrim.pyx:71-72 */
    *start_out = start; *stop_out = stop;
}

/* nextseq_trim_index (qualtrim.pyx:76-117): like the 3' pass above, but a 'G' counts as quality
   cutoff - 1.  Returns the index at which the read is cut. */
int oracle_nextseq_trim_index(const unsigned char *seq, const unsigned char *qual, int n, int cutoff,
                              int base)
{
    int s = 0, max_qual = 0, max_i = n, i;
    for (i = n - 1; i >= 0; i--) {                                /* qualtrim.pyx:106-116 */
        int q = (signed char)qual[i] - base;
        if (seq[i] == 'G') q = cutoff - 1;
        s += cutoff - q;
        if (s < 0) break;
        if (s > max_qual) { max_qual = s; max_i = i; }
    }
    return max_i;
}

/* poly_a_trim_index (qualtrim.pyx:120-169): start of the poly-A tail, or with revcomp the end of
   the poly-T head; +1 per A (T), -2 otherwise, at most 20 % errors, tails shorter than 3 ignored. */
int oracle_poly_a_trim_index(const unsigned char *seq, int n, int revcomp)
{
    int best_score = 0, score = 0, errors = 0, best_index, i;
    if (revcomp) {
        best_index = 0;
        for (i = 0; i < n; i++) {                                 /* qualtrim.pyx:141-153 */
            if (seq[i] == 'T') score += 1; else { score -= 2; errors += 1; }
            if (score > best_score && errors * 5 <= i + 1) { best_score = score; best_index = i + 1; }
        }
        if (best_index < 3) best_index = 0;
    } else {
        best_index = n;
        for (i = n - 1; i >= 0; i--) {                            /* qualtrim.pyx:155-167 */
            if (seq[i] == 'A') score += 1; else { score -= 2; errors += 1; }
            if (score > best_score && errors * 5 <= n - i) { best_score = score; best_index = i; }
        }
        if (best_index > n - 3) best_index = n;
    }
    return best_index;
}

/* expected_errors (qualtrim.pyx:172-197; expected_errors.h:95-140): sum of 10^(-q/10) over the
   qualities, in four independent accumulators over groups of four characters plus a tail that goes
   into the first, added up as ((e0 + e1) + e2) + e3.  The order is part of the result (FP64).
   Returns -1.0 for a character outside [base, 126].  The reference's 94-entry table holds the
   doubles nearest to 10^(-q/10); pow() reproduces all of them (checked by the golden test). */
double oracle_expected_errors(const unsigned char *qual, size_t n, unsigned char base)
{
    static double table[256];
    static int ready = 0;
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
    unsigned char max_phred = (unsigned char)(126 - base);
    size_t i = 0;
    if (!ready) {
        int q;
        for (q = 0; q < 256; q++) table[q] = pow(10.0, -(double)q / 10.0);
        ready = 1;
    }
    while (i + 3 < n) {
        unsigned char p0 = (unsigned char)(qual[i] - base), p1 = (unsigned char)(qual[i + 1] - base);
        unsigned char p2 = (unsigned char)(qual[i + 2] - base), p3 = (unsigned char)(qual[i + 3] - base);
        if (p0 > max_phred || p1 > max_phred || p2 > max_phred || p3 > max_phred) return -1.0;
        e0 += table[p0]; e1 += table[p1]; e2 += table[p2]; e3 += table[p3];
        i += 4;
    }
    while (i < n) {
        unsigned char ph = (unsigned char)(qual[i] - base);
        if (ph > max_phred) return -1.0;
        e0 += table[ph];
        i += 1;
    }
    return e0 + e1 + e2 + e3;
}

/* ------------------------------------------------------------------------------------ */
/* Batch drivers used by bench.py's "port" CPU baseline and by the tests                 */
/* ------------------------------------------------------------------------------------ */

/* locate() over a packed batch: seq bytes + n+1 offsets; results 8 x int32 per read:
   {found, ref_start, ref_stop, query_start, query_stop, score, errors, 0} */
int oracle_locate_batch(const unsigned char *ref, int m, const unsigned char *seq,
                        const int64_t *offsets, int64_t n_reads, double rate, int flags,
                        int wildcard_ref, int wildcard_query, int indel_cost, int min_overlap,
                        const oracle_kmer_entry *entries, const uint64_t *masks, int n_entries,
                        int32_t *results)
{
    int64_t r;
    for (r = 0; r < n_reads; r++) {
        const unsigned char *q = seq + offsets[r];
        int n = (int)(offsets[r + 1] - offsets[r]), out[6] = {0, 0, 0, 0, 0, 0}, rc = 1;
        int32_t *res = results + 8 * r;
        if (n_entries > 0) rc = oracle_kmers_present(entries, masks, n_entries, q, n);
        if (rc == 1)
            rc = oracle_locate(ref, m, q, n, rate, flags, wildcard_ref, wildcard_query,
                               indel_cost, min_overlap, out);
        if (rc < 0) return rc;
        res[0] = rc; res[1] = out[0]; res[2] = out[1]; res[3] = out[2]; res[4] = out[3];
        res[5] = out[4]; res[6] = out[5]; res[7] = 0;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* The whole per-read pass in C (the composition oracle.py restates in Python), so that   */
/* the parity tests can check >= 10^6 reads per configuration in seconds; oracle.py runs  */
/* it from several threads over disjoint read ranges.                                     */
/*   <Adapter>.match_to        adapters.py:707-724, 758-786, 815-832, 862-890, 915-935,   */
/*                             963-975, 1000-1012                                         */
/*   LinkedAdapter.match_to    adapters.py:1215-1227 (score/errors: 1113-1130)            */
/*   MultipleAdapters.match_to adapters.py:1265-1286                                      */
/*   NextseqQualityTrimmer, QualityTrimmer, AdapterCutter rounds                          */
/*                             modifiers.py:825-858, 225-231 (order: cli.py:940-953)      */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const unsigned char *seq;   /* adapter characters as the aligner gets them */
    int32_t m;
    int32_t flags, wildcard_ref, wildcard_query, indel_cost, min_overlap;
    int32_t kind;               /* 0 aligner, 1 PrefixComparer, 2 SuffixComparer */
    int32_t reverse_read;       /* Rightmost* adapters: search the reversed read (adapters.py:766, 870) */
    int32_t remove;             /* 0 before, 1 after, 2 auto (AnywhereAdapter, adapters.py:930-935) */
    int32_t n_entries;          /* 0: no KmerFinder (MockKmerFinder: always present) */
    double max_error_rate;
    const oracle_kmer_entry *entries;
    const uint64_t *masks;
} oracle_adapter;

typedef struct { int32_t type, a0, a1, front_required, back_required; } oracle_group;   /* type 0 single, 1 linked */

typedef struct { int adapter, astart, astop, rstart, rstop, score, errors, remove; } oracle_hit;

/* returns 1 match, 0 none, < 0 error */
static int batch_match_single(const oracle_adapter *A, int ai, const unsigned char *s, int n, unsigned char *tmp,
                              oracle_hit *h)
{
    const unsigned char *q = s;
    int out[6], rc, i;
    if (A->reverse_read) {
        for (i = 0; i < n; i++) tmp[i] = s[n - 1 - i];
        q = tmp;
    }
    if (A->n_entries > 0) {
        rc = oracle_kmers_present(A->entries, A->masks, A->n_entries, q, n);
        if (rc <= 0) return rc;
    }
    if (A->kind == 0)
        rc = oracle_locate(A->seq, A->m, q, n, A->max_error_rate, A->flags, A->wildcard_ref, A->wildcard_query,
                           A->indel_cost, A->min_overlap, out);
    else if (A->kind == 1)
        rc = oracle_prefix_compare(A->seq, A->m, q, n, A->max_error_rate, A->wildcard_ref, A->wildcard_query,
                                   A->min_overlap, out);
    else
        rc = oracle_suffix_compare(A->seq, A->m, q, n, A->max_error_rate, A->wildcard_ref, A->wildcard_query,
                                   A->min_overlap, out);
    if (rc <= 0) return rc;
    h->adapter = ai;
    h->astart = out[0]; h->astop = out[1]; h->rstart = out[2]; h->rstop = out[3];
    h->score = out[4]; h->errors = out[5];
    if (A->reverse_read) {                                        /* adapters.py:777-785, 881-889 */
        h->astart = A->m - out[1]; h->astop = A->m - out[0];
        h->rstart = n - out[3]; h->rstop = n - out[2];
    }
    h->remove = A->remove;
    if (h->remove == 2) h->remove = h->rstart == 0 ? 0 : 1;       /* adapters.py:930-935 */
    return 1;
}

static void batch_trim(const oracle_hit *h, int *s, int *e)       /* Match.trimmed: adapters.py:453-454, 486-487 */
{
    if (h->remove == 0) *s += h->rstop; else *e = *s + h->rstart;
}

static void batch_put(int32_t *rec, const oracle_hit *h, int group, int searched)
{
    rec[0] = h->adapter; rec[1] = h->astart; rec[2] = h->astop; rec[3] = h->rstart; rec[4] = h->rstop;
    rec[5] = h->score; rec[6] = h->errors;
    rec[7] = (int32_t)((uint32_t)(group & 255) | (h->remove == 1 ? 256u : 0u) | ((uint32_t)(searched & 0xFFFF) << 16));
}

/*
 * records: (r_end - r_begin) x times x slots x 8 int32 in the cg_match layout (adapter = -1: none),
 * qtrim: (r_end - r_begin) x 2.  Both are indexed relative to r_begin.  max_len bounds the reads.
 */
int oracle_process_batch(const oracle_adapter *ads, int n_ads, const oracle_group *groups, int n_groups,
                         const unsigned char *seq, const unsigned char *qual, const int64_t *offsets,
                         int64_t r_begin, int64_t r_end, int quality_trim, int cutoff_front, int cutoff_back,
                         int qbase, int times, int use_nextseq, int nextseq_cutoff, int slots,
                         int32_t *records, int32_t *qtrim)
{
    int64_t r;
    int max_len = 0, rc = 0;
    unsigned char *tmp;
    (void)n_ads;
    for (r = r_begin; r < r_end; r++)
        if (offsets[r + 1] - offsets[r] > max_len) max_len = (int)(offsets[r + 1] - offsets[r]);
    tmp = malloc((size_t)max_len + 1);
    if (!tmp) return -4;
    for (r = r_begin; r < r_end && rc >= 0; r++) {
        const unsigned char *sq = seq + offsets[r];
        const int n = (int)(offsets[r + 1] - offsets[r]);
        int32_t *rec = records + (size_t)(r - r_begin) * times * slots * 8;
        int s = 0, e = n, round, i;
        for (i = 0; i < times * slots; i++) { memset(rec + 8 * i, 0, 32); rec[8 * i] = -1; }
        if (use_nextseq) e = oracle_nextseq_trim_index(sq, qual + offsets[r], n, nextseq_cutoff, qbase);
        if (quality_trim) oracle_quality_trim_index(qual + offsets[r], e, cutoff_front, cutoff_back, qbase, &s, &e);
        qtrim[2 * (r - r_begin)] = s; qtrim[2 * (r - r_begin) + 1] = e;
        for (round = 0; round < times; round++) {                 /* modifiers.py:225-231 */
            int have = 0, best_group = -1, best_score = 0, best_errors = 0, g;
            oracle_hit b0, b1;
            b0.adapter = -1; b1.adapter = -1;
            for (g = 0; g < n_groups; g++) {                      /* adapters.py:1271-1286 */
                const oracle_group *G = &groups[g];
                oracle_hit h0, h1;
                int score, errors, f, bk;
                h0.adapter = -1; h1.adapter = -1;
                if (G->type == 0) {
                    f = batch_match_single(&ads[G->a0], G->a0, sq + s, e - s, tmp, &h0);
                    if (f < 0) { rc = f; break; }
                    if (!f) continue;
                    score = h0.score; errors = h0.errors;
                } else {                                          /* adapters.py:1215-1227 */
                    int s2 = 0, e2 = e - s;
                    f = batch_match_single(&ads[G->a0], G->a0, sq + s, e - s, tmp, &h0);
                    if (f < 0) { rc = f; break; }
                    if (!f) h0.adapter = -1;
                    if (G->front_required && !f) continue;
                    if (f) batch_trim(&h0, &s2, &e2);
                    bk = batch_match_single(&ads[G->a1], G->a1, sq + s + s2, e2 - s2, tmp, &h1);
                    if (bk < 0) { rc = bk; break; }
                    if (!bk) h1.adapter = -1;
                    if (!bk && (G->back_required || !f)) continue;
                    score = (f ? h0.score : 0) + (bk ? h1.score : 0);   /* adapters.py:1113-1130 */
                    errors = (f ? h0.errors : 0) + (bk ? h1.errors : 0);
                }
                if (!have || score > best_score || (score == best_score && errors < best_errors)) {
                    have = 1; best_score = score; best_errors = errors; best_group = g; b0 = h0; b1 = h1;
                }
            }
            if (rc < 0 || !have) break;                           /* modifiers.py:227-229 */
            {
                const int searched = e - s;
                int32_t *dst = rec + (size_t)round * slots * 8;
                if (b0.adapter >= 0) batch_put(dst, &b0, best_group, searched);
                if (b1.adapter >= 0 && slots > 1) {
                    int s2 = 0, e2 = searched;
                    if (b0.adapter >= 0) batch_trim(&b0, &s2, &e2);
                    batch_put(dst + 8, &b1, best_group, e2 - s2);
                }
                if (b0.adapter >= 0) batch_trim(&b0, &s, &e);      /* modifiers.py:231 */
                if (b1.adapter >= 0) batch_trim(&b1, &s, &e);
            }
        }
    }
    free(tmp);
    return rc < 0 ? rc : 0;
}
