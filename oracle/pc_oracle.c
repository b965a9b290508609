/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the prover side of Virgo's polynomial commitment:
 * poly_commit_prover::commit_private_array, commit_public_array and the FRI commit phase (fri::commit_phase_step).
 * Plain-C restatement; every function cites the reference lines it follows
 * (paths under /root/reference/lib/virgo/src). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this; the product never does.
 *
 * Parity status: PINNED -- tests/test_oracle.py compares it with the UNMODIFIED reference (oracle/_ref/ref_pc_commit,
 * the reference's own poly_commit_prover linked with its prebuilt XKCP SHA3) through the golden files
 * tests/golden/pc_*.json written by tests/golden/make_golden_pc*.py: Merkle roots, SHA-256 of the codeword arrays, of the
 * leaf hashes and Merkle trees, all_sum, the virtual oracle, every FRI level's root / codewords / tree.
 *
 * Third-party code on this path that is NOT in source form under /root/reference: SHA3-256 from the Keccak team's
 * XKCP (prebuilt libXKCP.a, version unrecorded; header lib/libXKCP.a.headers/SimpleFIPS202.h), called only through
 * my_hhash.h:27-33 on 64-byte messages. Restated here from the published algorithm (FIPS 202: Keccak-f[1600], rate
 * 136 bytes, domain suffix 0x06, one permutation per 64-byte message) and pinned by the FIPS 202 known answers
 * (tests/test_oracle.py) plus the reference's own outputs above.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gkr_oracle.h"

typedef uint64_t u64;

/* ------------------------------------------------------------------ SHA3-256 (FIPS 202) */
static const u64 KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
static inline u64 rotl64(u64 x, int n) { return (x << n) | (x >> (64 - n)); }
static void keccak_f1600(u64 st[25]) {
    for (int round = 0; round < 24; ++round) {
        u64 bc[5];
        for (int i = 0; i < 5; ++i) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; ++i) {
            u64 t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        u64 t = st[1];
        for (int i = 0; i < 24; ++i) {
            int j = KECCAK_PIL[i];
            u64 b = st[j];
            st[j] = rotl64(t, KECCAK_ROT[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; ++i) bc[i] = st[j + i];
            for (int i = 0; i < 5; ++i) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KECCAK_RC[round];
    }
}
/* general-length SHA3-256 (used by the known-answer tests); the commitment only hashes 64-byte blocks */
void opc_sha3_256(const unsigned char* msg, size_t len, unsigned char out[32]) {
    u64 st[25];
    memset(st, 0, sizeof st);
    const size_t rate = 136;
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; ++i) {
            u64 w;
            memcpy(&w, msg + 8 * i, 8);
            st[i] ^= w;
        }
        keccak_f1600(st);
        msg += rate;
        len -= rate;
    }
    unsigned char blk[136];
    memset(blk, 0, sizeof blk);
    memcpy(blk, msg, len);
    blk[len] ^= 0x06;
    blk[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; ++i) {
        u64 w;
        memcpy(&w, blk + 8 * i, 8);
        st[i] ^= w;
    }
    keccak_f1600(st);
    memcpy(out, st, 32);
}
/* my_hhash.h:27-33: SHA3_256(dst, src, 64) */
static void hhash64(const void* src, void* dst) { opc_sha3_256((const unsigned char*)src, 64, (unsigned char*)dst); }

/* ------------------------------------------------------------------ field helpers */
static const ofe ONE = {1, 0}, ZERO = {0, 0};
static ofe f_pow(ofe x, unsigned __int128 e) { /* fieldElement.cpp:322-334 fastPow */
    ofe ret = ONE, tmp = x;
    while (e) {
        if (e & 1) ret = ofe_mul(ret, tmp);
        tmp = ofe_mul(tmp, tmp);
        e >>= 1;
    }
    return ret;
}
/* fieldElement.cpp:237-249 getRootOfUnity: the order-2^62 element squared (62 - log_order) times */
ofe opc_root_of_unity(int log_order) {
    ofe rou = {2147483648ULL, 1033321771269002680ULL};
    for (int i = 0; i < 62 - log_order; ++i) rou = ofe_mul(rou, rou);
    return rou;
}

/* result[k] = sum_i coef[i] * w^(i*k), k < order, natural order; coef_len <= order, both powers of two.
 * RS_polynomial.cpp:26-153 computes exactly this (its packed / unrolled stages are an implementation of the same
 * radix-2 decimation in time; field arithmetic is exact, so any evaluation order gives the same canonical values). */
static void fft(const ofe* coef, int coef_len, int order, ofe w, ofe* result) {
    /* iterative radix-2 DIT on a zero-padded copy */
    int lg = 0;
    while ((1 << lg) < order) ++lg;
    ofe* a = (ofe*)calloc((size_t)order, sizeof(ofe));
    for (int i = 0; i < order; ++i) {  /* bit reversal */
        unsigned r = 0;
        for (int b = 0; b < lg; ++b) r |= ((unsigned)(i >> b) & 1u) << (lg - 1 - b);
        a[r] = i < coef_len ? coef[i] : ZERO;
    }
    for (int s = 1; s <= lg; ++s) {
        const int m = 1 << s, half = m >> 1;
        ofe wm = w;
        for (int i = 0; i < lg - s; ++i) wm = ofe_mul(wm, wm);  /* w^(order/m) */
        for (int k = 0; k < order; k += m) {
            ofe x = ONE;
            for (int j = 0; j < half; ++j) {
                ofe t = ofe_mul(x, a[k + j + half]), u = a[k + j];
                a[k + j] = ofe_add(u, t);
                a[k + j + half] = ofe_sub(u, t);
                x = ofe_mul(x, wm);
            }
        }
    }
    memcpy(result, a, (size_t)order * sizeof(ofe));
    free(a);
}
/* RS_polynomial.cpp:155-220 inverse_fast_fourier_transform with coef_len == order (the only use on this path):
 * fft with the inverse root, then times order^(p-2) */
static void ifft(const ofe* evals, int order, ofe root, ofe* dst) {
    int lg = 0;
    while ((1 << lg) < order) ++lg;
    ofe inv_rou = ONE, tmp = root;
    for (int i = 0; i < lg; ++i) {
        inv_rou = ofe_mul(inv_rou, tmp);
        tmp = ofe_mul(tmp, tmp);
    }
    fft(evals, order, order, inv_rou, dst);
    ofe n = {(u64)order, 0};
    ofe inv_n = f_pow(n, (unsigned __int128)2305843009213693951ULL - 2);
    for (int i = 0; i < order; ++i) dst[i] = ofe_mul(dst[i], inv_n);
}

/* ------------------------------------------------------------------ commit_private_array
 * poly_commit.h:41-124 + vpd_prover.cpp:9-14 + fri.cpp:36-139 (request_init_commit, oracle 0) + merkle_tree.cpp:7-51.
 * constants.h: log_slice_number = 6, rs_code_rate = 5.
 *   array: 2^log_len elements; mask: n_mask elements (the GKR prover passes one zero, prover.cpp:524-530).
 * Outputs (any may be NULL): l_eval [65 * slice_size], leaf_hash [slice_size/2 * 32 B], tree [slice_size * 32 B]
 * (array heap: node 1 = root, leaves at [slice_size/2, slice_size)), root [32 B]. slice_size = 2^(log_len - 1).
 * Returns slice_size, or -1 if log_len < 6. */
long opc_commit_private(const ofe* array, int log_len, const ofe* mask, int n_mask, ofe* l_eval_out, unsigned char* leaf_out,
                        unsigned char* tree_out, unsigned char root_out[32]) {
    const int LOG_SLICE = 6, RATE = 5, SLICES = 1 << LOG_SLICE;
    if (log_len < LOG_SLICE || n_mask < 1) return -1;
    const int slice_count = SLICES + 1;
    const int slice_size = 1 << (log_len + RATE - LOG_SLICE);
    const int real_cnt = slice_size >> RATE;
    ofe* l_eval = (ofe*)calloc((size_t)slice_count * slice_size, sizeof(ofe));
    /* mask placement, poly_commit.h:55-67 */
    int gap = slice_size / n_mask;
    for (int j = 0; j < 31; ++j)
        if ((1 << j) <= gap && (1 << (j + 1)) > gap) { gap = 1 << j; break; }
    const int mask_n = slice_size / gap;
    ofe* tmp = (ofe*)calloc((size_t)(mask_n > real_cnt ? mask_n : real_cnt), sizeof(ofe));
    ofe* m = (ofe*)calloc((size_t)mask_n, sizeof(ofe));
    for (int j = 0; j < n_mask && j < mask_n; ++j) m[j] = mask[j];
    int lg_mask = 0, lg_real = 0, lg_slice = 0;
    while ((1 << lg_mask) < mask_n) ++lg_mask;
    while ((1 << lg_real) < real_cnt) ++lg_real;
    while ((1 << lg_slice) < slice_size) ++lg_slice;
    for (int i = 0; i < slice_count; ++i) {
        if (i == slice_count - 1) {
            ifft(m, mask_n, opc_root_of_unity(lg_mask), tmp);
            fft(tmp, mask_n, slice_size, opc_root_of_unity(lg_slice), l_eval + (size_t)i * slice_size);
        } else {   /* (an all-zero slice short-cuts to zeros in the reference: same values) */
            ifft(array + (size_t)i * real_cnt, real_cnt, opc_root_of_unity(lg_real), tmp);
            fft(tmp, real_cnt, slice_size, opc_root_of_unity(lg_slice), l_eval + (size_t)i * slice_size);
        }
    }
    /* fri.cpp:84-126: leaf i = chain over the 64 slices of H(eval_s[i] || eval_s[i + half] || previous), then the mask slice */
    const int half = slice_size / 2;
    unsigned char* leaf = (unsigned char*)calloc((size_t)half, 32);
    for (int i = 0; i < half; ++i) {
        unsigned char h[32], data[64];
        memset(h, 0, 32);
        for (int s = 0; s < slice_count; ++s) {
            memcpy(data, &l_eval[(size_t)s * slice_size + i], 16);
            memcpy(data + 16, &l_eval[(size_t)s * slice_size + i + half], 16);
            memcpy(data + 32, h, 32);
            hhash64(data, h);
        }
        memcpy(leaf + (size_t)i * 32, h, 32);
    }
    /* merkle_tree.cpp:7-51 (ele_num = half is a power of two: no padding leaves) */
    unsigned char* tree = (unsigned char*)calloc((size_t)slice_size, 32);
    memcpy(tree + (size_t)half * 32, leaf, (size_t)half * 32);
    for (int lvl = half / 2; lvl >= 1; lvl /= 2)
        for (int i = 0; i < lvl; ++i) hhash64(tree + (size_t)(2 * (lvl + i)) * 32, tree + (size_t)(lvl + i) * 32);
    if (half == 1) { /* single leaf: create_tree leaves node 1 = the leaf itself */ }
    if (root_out) memcpy(root_out, tree + 32, 32);
    if (l_eval_out) memcpy(l_eval_out, l_eval, (size_t)slice_count * slice_size * sizeof(ofe));
    if (leaf_out) memcpy(leaf_out, leaf, (size_t)half * 32);
    if (tree_out) memcpy(tree_out, tree, (size_t)slice_size * 32);
    free(l_eval); free(tmp); free(m); free(leaf); free(tree);
    return slice_size;
}

/* ------------------------------------------------------------------ commit_public_array
 * poly_commit.h:126-349 for the GKR use of it (verifier.cpp:367-383: public array = the eq table over r_liu, ONE zero
 * public mask, private mask one zero) + fri::request_init_commit(.., 1) (fri.cpp:36-139) on h_eval_arr.
 *   q_eval: the public array encoded like the private one (per-slice inverse FFT + 32x extension), mask slice zero;
 *   per slice: the 2n evaluations of l*q on the 2n-th roots (every 16th codeword position) -> 2n coefficients; the upper n
 *   are h (l*q = g + (x^n - 1) h), evaluated on all N points -> h_eval; all_sum = n (lq_0 + h_0);
 *   virtual_oracle_witness[(j mod N/2) << 7 | slice << 1 | (j >= N/2)] = (l q - (x^n - 1) h - (lq_0 + h_0)) * n * x^-1, x = w_N^j.
 * With zero masks the 65th slice of everything is zero.
 * Outputs (any may be NULL): all_sum [65], h_eval [65 N], vow [64 N], root_h [32]. Returns N, or -1. */
long opc_commit_public(const ofe* array, const ofe* pub, int log_len, ofe* all_sum_out, ofe* h_eval_out, ofe* vow_out,
                       unsigned char root_out[32]) {
    const int LOG_SLICE = 6, RATE = 5, SLICES = 1 << LOG_SLICE;
    if (log_len < LOG_SLICE) return -1;
    const int slice_count = SLICES + 1;
    const int N = 1 << (log_len + RATE - LOG_SLICE), n = N >> RATE, half = N / 2;
    int lg_n = 0, lg_N = 0, lg_2n = 0;
    while ((1 << lg_n) < n) ++lg_n;
    while ((1 << lg_N) < N) ++lg_N;
    while ((1 << lg_2n) < 2 * n) ++lg_2n;
    ofe* l_eval = (ofe*)calloc((size_t)slice_count * N, sizeof(ofe));
    ofe* q_eval = (ofe*)calloc((size_t)slice_count * N, sizeof(ofe));
    ofe* h_eval = (ofe*)calloc((size_t)slice_count * N, sizeof(ofe));
    ofe* vow = (ofe*)calloc((size_t)SLICES * N, sizeof(ofe));
    ofe* tmp = (ofe*)calloc((size_t)2 * n, sizeof(ofe));
    ofe* lq = (ofe*)calloc((size_t)2 * n, sizeof(ofe));
    ofe all_sum[65];
    for (int i = 0; i < SLICES; ++i) {
        ifft(array + (size_t)i * n, n, opc_root_of_unity(lg_n), tmp);
        fft(tmp, n, N, opc_root_of_unity(lg_N), l_eval + (size_t)i * N);
        ifft(pub + (size_t)i * n, n, opc_root_of_unity(lg_n), tmp);
        fft(tmp, n, N, opc_root_of_unity(lg_N), q_eval + (size_t)i * N);
    }
    const ofe rou = opc_root_of_unity(lg_N);
    ofe inv_rou = ONE;   /* rou^(N-1) */
    {
        ofe t = rou;
        for (int b = 0; b < lg_N; ++b) { inv_rou = ofe_mul(inv_rou, t); t = ofe_mul(t, t); }
    }
    const ofe rou_n = f_pow(rou, (unsigned __int128)n);
    const ofe n_fe = {(u64)n, 0};
    for (int i = 0; i < SLICES; ++i) {
        const int step = N / (2 * n);
        for (int j = 0; j < 2 * n; ++j) lq[j] = ofe_mul(l_eval[(size_t)i * N + (size_t)j * step], q_eval[(size_t)i * N + (size_t)j * step]);
        ifft(lq, 2 * n, opc_root_of_unity(lg_2n), tmp);          /* tmp = lq_coef */
        fft(tmp + n, n, N, rou, h_eval + (size_t)i * N);            /* h_coef = upper half */
        const ofe c0 = ofe_add(tmp[0], tmp[n]);
        all_sum[i] = ofe_mul(c0, n_fe);
        const ofe const_sum = ofe_sub(ZERO, c0);
        ofe inv_x = n_fe, x_n = ONE;
        for (int j = 0; j < N; ++j) {
            const ofe lqv = ofe_mul(l_eval[(size_t)i * N + j], q_eval[(size_t)i * N + j]);
            const ofe g = ofe_sub(lqv, ofe_mul(ofe_sub(x_n, ONE), h_eval[(size_t)i * N + j]));
            const ofe v = ofe_mul(ofe_add(g, const_sum), inv_x);
            if (j < half) vow[((size_t)j << (LOG_SLICE + 1)) | ((size_t)i << 1)] = v;
            else vow[((size_t)(j - half) << (LOG_SLICE + 1)) | ((size_t)i << 1) | 1] = v;
            inv_x = ofe_mul(inv_x, inv_rou);
            x_n = ofe_mul(x_n, rou_n);
        }
    }
    all_sum[SLICES] = ZERO;
    /* Merkle commitment of h_eval_arr (same leaf chain and tree as for l_eval) */
    unsigned char* tree = (unsigned char*)calloc((size_t)N, 32);
    for (int i = 0; i < half; ++i) {
        unsigned char h[32], data[64];
        memset(h, 0, 32);
        for (int s2 = 0; s2 < slice_count; ++s2) {
            memcpy(data, &h_eval[(size_t)s2 * N + i], 16);
            memcpy(data + 16, &h_eval[(size_t)s2 * N + i + half], 16);
            memcpy(data + 32, h, 32);
            hhash64(data, h);
        }
        memcpy(tree + (size_t)(half + i) * 32, h, 32);
    }
    for (int lvl = half / 2; lvl >= 1; lvl /= 2)
        for (int i = 0; i < lvl; ++i) hhash64(tree + (size_t)(2 * (lvl + i)) * 32, tree + (size_t)(lvl + i) * 32);
    if (root_out) memcpy(root_out, tree + 32, 32);
    if (all_sum_out) memcpy(all_sum_out, all_sum, sizeof all_sum);
    if (h_eval_out) memcpy(h_eval_out, h_eval, (size_t)slice_count * N * sizeof(ofe));
    if (vow_out) memcpy(vow_out, vow, (size_t)SLICES * N * sizeof(ofe));
    free(l_eval); free(q_eval); free(h_eval); free(vow); free(tmp); free(lq); free(tree);
    return N;
}

/* ------------------------------------------------------------------ FRI commit phase
 * fri::commit_phase_step (fri.cpp:289-418) as poly_commit_prover::commit_phase drives it (vpd_verifier.cpp:43-73: one step
 * per randomness until 32 points per slice are left), starting from the virtual oracle of commit_public_array, with the
 * zero mask codeword of the GKR use (rs_codeword_msk stays zero).
 * A level holds, per slice j, a codeword f_j of M points on the M-th roots, stored as pairs of opposite points:
 *   f_j[k] at  (k mod M/2) << 7 | j << 1 | (k >= M/2)          (virtual_oracle_witness_mapping / rs_codeword_mapping)
 * one step: g_j[i] = ( f_j[i] + f_j[i + M/2]  +  r w_M^-i ( f_j[i] - f_j[i + M/2] ) ) / 2,  i < M/2   (:316-337; pos == i always:
 * (M/2 + i)/2 >= i for i < M/2), stored the same way with M/2 for M (:339-357); leaf i < M/4 = the SHA3 chain over the
 * 64 pairs of leaf i and then the (zero) mask pair (:383-404); array-heap tree over the M/4 leaves (:405).
 *   vow: 64 N elements; randomness: n_steps elements, n_steps <= log_N - 5.
 * Outputs (any may be NULL): roots [n_steps * 32]; codes: the levels back to back, level l = 64 * (N >> (l+1)) elements;
 * trees: level l = (N >> (l+1)) * 32 bytes (node 0 zero, node 1 the root). Returns the number of steps done, or -1. */
long opc_fri_commit_phase(const ofe* vow, int log_N, const ofe* randomness, int n_steps, unsigned char* roots_out, ofe* codes_out,
                          unsigned char* trees_out) {
    const int LOG_SLICE = 6, RATE = 5, SLICES = 1 << LOG_SLICE;
    if (log_N < RATE || n_steps < 0 || n_steps > log_N - RATE) return -1;
    const size_t N = (size_t)1 << log_N;
    const ofe two = {2, 0};
    const ofe inv2 = f_pow(two, (unsigned __int128)2305843009213693951ULL - 2);
    ofe* prev = (ofe*)malloc((size_t)SLICES * N * sizeof(ofe));
    memcpy(prev, vow, (size_t)SLICES * N * sizeof(ofe));
    size_t code_off = 0, tree_off = 0;
    for (int s = 0; s < n_steps; ++s) {
        const size_t M = N >> s, half = M / 2, quarter = M / 4;
        const ofe r = randomness[s];
        const ofe w = opc_root_of_unity(log_N - s);
        const ofe w_inv = f_pow(w, (unsigned __int128)(M - 1));
        ofe* cur = (ofe*)malloc((size_t)SLICES * half * sizeof(ofe));
        ofe inv_mu = ONE;   /* w_M^-i */
        for (size_t i = 0; i < half; ++i) {
            const ofe rm = ofe_mul(inv_mu, r);
            for (int j = 0; j < SLICES; ++j) {
                const ofe a = prev[(i << (LOG_SLICE + 1)) | ((size_t)j << 1)], b = prev[(i << (LOG_SLICE + 1)) | ((size_t)j << 1) | 1];
                const ofe v = ofe_mul(inv2, ofe_add(ofe_add(a, b), ofe_mul(rm, ofe_sub(a, b))));
                cur[((i % quarter) << (LOG_SLICE + 1)) | ((size_t)j << 1) | (i >= quarter)] = v;
            }
            inv_mu = ofe_mul(inv_mu, w_inv);
        }
        unsigned char* tree = (unsigned char*)calloc(half, 32);
        for (size_t i = 0; i < quarter; ++i) {
            unsigned char h[32], data[64];
            memset(h, 0, 32);
            for (int j = 0; j <= SLICES; ++j) {
                if (j < SLICES) memcpy(data, &cur[(i << (LOG_SLICE + 1)) | ((size_t)j << 1)], 32);
                else memset(data, 0, 32);   /* rs_codeword_msk: zero */
                memcpy(data + 32, h, 32);
                hhash64(data, h);
            }
            memcpy(tree + (quarter + i) * 32, h, 32);
        }
        for (size_t lvl = quarter / 2; lvl >= 1; lvl /= 2)
            for (size_t i = 0; i < lvl; ++i) hhash64(tree + (2 * (lvl + i)) * 32, tree + (lvl + i) * 32);
        if (roots_out) memcpy(roots_out + (size_t)s * 32, tree + 32, 32);
        if (codes_out) memcpy(codes_out + code_off, cur, (size_t)SLICES * half * sizeof(ofe));
        if (trees_out) memcpy(trees_out + tree_off, tree, half * 32);
        code_off += (size_t)SLICES * half;
        tree_off += half * 32;
        free(tree);
        free(prev);
        prev = cur;
    }
    free(prev);
    return n_steps;
}
