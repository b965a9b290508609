// The polynomial commitment's inner GKR on the device -- included by engine.cu (it drives the stand-alone
// sumcheck objects defined there).
// Replaces lib/virgo/src/fft_circuit_GKR.cpp: fft_gkr (:833-849) = build_circuit (:21-101) + engage_gkr (:784-831):
// a prover and a verifier run a layered GKR on a fixed circuit family --
//   E    the eq table of a random point r (2^lg values)                                   :24-32
//   F_d  lg inverse-FFT butterfly layers, d = lg-1 .. 0                                   :34-65
//   S    F_0 / 2^lg                                                                       :66-71
//   P    64 x 2^lg products S[j] x_i^j for 64 random points x_i                           :73-90
//   O    their 64 row sums (the polynomial with coefficients S evaluated at the x_i)      :91-100
// -- from O back to E: one sumcheck for O and for P (addition_layer :224-331, mult_layer :333-445), a rescaling
// (intermediate_layer :447-456) and two sumchecks per butterfly layer (ifft_gkr :458-769: phase 1 binds the first operand,
// phase 2 the second). Every sumcheck has the shape sum_i M[i] V[i] + A[i] of the GKR prover's own rounds (sumcheck_phase1/2_
// update :155-222 == prover.cpp:457-492), so each one is ONE launch of the pass kernel (k_phase_dfs, two rounds per pass).
// All randomness comes from fieldElement::random() and never depends on a prover message: the caller hands it over as one
// array in the reference's draw order, the device runs the whole prover side without a host round trip (the only value
// that feeds back, v_u into phase 2's tables, is read on the device), and the verifier's closed-form checks run on the
// host afterwards.
#pragma once

namespace fg {

struct Pt { F v[30]; };   // a random point (lg + 6 <= 30 coordinates), passed by value

VP_D F eq_pt(const Pt& r, int bits, uint32_t g) {   // prod_k (g_k ? r_k : 1 - r_k)
    F a = f_one();
    for (int k = 0; k < bits; ++k) a = f_mul(a, ((g >> k) & 1u) ? r.v[k] : f_sub(f_one(), r.v[k]));
    return a;
}
VP_D F pow_bits(const F* __restrict__ pw, int bits, uint32_t e) {   // x^e from pw[b] = x^(2^b)
    F a = f_one();
    for (int b = 0; b < bits; ++b)
        if ((e >> b) & 1u) a = f_mul(a, pw[b]);
    return a;
}
VP_D F ldg(const F* p) {
    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p);
    return F{t.x, t.y};
}
VP_D void stg(F* p, const F& v) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(v.re, v.im); }

// E: level i of build_circuit maps j -> (2j: times r_i, 2j + 1: times 1 - r_i), so bit (lg-1-i) of the final index picks
__global__ void k_fg_eq_layer(F* __restrict__ out, int lg, Pt r) {
    const uint32_t n = 1u << lg;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        F a = f_one();
        for (int i = 0; i < lg; ++i) a = f_mul(a, ((g >> (lg - 1 - i)) & 1u) ? f_sub(f_one(), r.v[i]) : r.v[i]);
        stg(out + g, a);
    }
}
// tw[t] = inv_rou^t, t < 2^lg, from sq[b] = inv_rou^(2^b)
__global__ void k_fg_twiddles(F* __restrict__ tw, int lg, Pt sq) {
    const uint32_t n = 1u << lg;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) stg(tw + t, pow_bits(sq.v, lg, t));
}
// one butterfly layer (:45-65): cur[k, j] = pre[k, 0, j] + x_k pre[k, 1, j], cur[k + hb, j] = pre[k, 0, j] - x_k pre[k, 1, j],
// x_k = (inv_rou^(2^dep))^k
__global__ void k_fg_butterfly(const F* __restrict__ pre, F* __restrict__ cur, int lg, int dep, const F* __restrict__ tw) {
    const uint32_t pairs = 1u << (lg - 1), cols = 1u << dep, hb = 1u << (lg - dep - 1);
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += gridDim.x * blockDim.x) {
        const uint32_t j = q & (cols - 1), k = q >> dep;
        const F l = ldg(pre + ((k << (dep + 1)) | j)), rv = f_mul(ldg(tw + (k << dep)), ldg(pre + ((k << (dep + 1)) | cols | j)));
        stg(cur + ((k << dep) | j), f_add(l, rv));
        stg(cur + (((k + hb) << dep) | j), f_sub(l, rv));
    }
}
__global__ void k_fg_scale(const F* __restrict__ in, F* __restrict__ out, F c, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) stg(out + i, f_mul(ldg(in + i), c));
}
// P[i][j] = S[j] * x_i^j (pw[i * 32 + b] = x_i^(2^b))
__global__ void k_fg_products(const F* __restrict__ S, const F* __restrict__ pw, int lg, F* __restrict__ P) {
    const size_t total = (size_t)64 << lg;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(w >> lg), j = (uint32_t)(w & ((1u << lg) - 1));
        stg(P + w, f_mul(ldg(S + j), pow_bits(pw + i * 32, lg, j)));
    }
}
// O[i] = sum_j P[i][j]: one block per i
__global__ void __launch_bounds__(256) k_fg_row_sums(const F* __restrict__ P, int lg, F* __restrict__ O) {
    __shared__ F sh[256];
    const uint32_t n = 1u << lg;
    const F* row = P + ((size_t)blockIdx.x << lg);
    F acc = f_zero();
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) acc = f_add(acc, ldg(row + j));
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = 128; s >= 1; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = f_add(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) stg(O + blockIdx.x, sh[0]);
}
// ---- sumcheck tables. addition_layer (:252-268): V = P, M[j] = alpha eq(r_0; row of j) + beta eq(r_1; row of j) (64 host values)
__global__ void k_fg_fill_add(const F* __restrict__ P, const F* __restrict__ row_w, int lg, F* __restrict__ V, F* __restrict__ M, F* __restrict__ A) {
    const size_t total = (size_t)64 << lg;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        stg(V + w, ldg(P + w));
        stg(M + w, ldg(row_w + (w >> lg)));
        stg(A + w, f_zero());
    }
}
// mult_layer (:360-383): V = S, M[i] = sum_j (alpha eq(r_0; j 2^lg + i) + beta eq(r_1; ..)) x_j^i. The eq of the lg + 6 bits
// splits into the part of i (low lg coordinates) and the part of j (hi0 / hi1: the 64 host values of the upper 6, times alpha / beta).
__global__ void k_fg_fill_mult(const F* __restrict__ S, Pt r0, Pt r1, const F* __restrict__ hi0, const F* __restrict__ hi1, int use1,
                               const F* __restrict__ pw, int lg, F* __restrict__ V, F* __restrict__ M, F* __restrict__ A) {
    const uint32_t n = 1u << lg;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        F s0 = f_zero(), s1 = f_zero();
        for (int j = 0; j < 64; ++j) {
            const F xp = pow_bits(pw + j * 32, lg, i);
            s0 = f_add(s0, f_mul(ldg(hi0 + j), xp));
            if (use1) s1 = f_add(s1, f_mul(ldg(hi1 + j), xp));
        }
        F m = f_mul(eq_pt(r0, lg, i), s0);
        if (use1) m = f_add(m, f_mul(eq_pt(r1, lg, i), s1));
        stg(V + i, ldg(S + i));
        stg(M + i, m);
        stg(A + i, f_zero());
    }
}
// butterfly layer, phase 1 (:524-560) and phase 2 (:585-610). g = (k, j) / (k + hb, j) differ in the top bit only:
//   t0 = alpha e0 (1 - r_0[lg-1]) + beta e1 (1 - r_1[lg-1]),  t1 = alpha e0 r_0[lg-1] + beta e1 r_1[lg-1],  e = eq over the low lg-1 bits
// phase 1: M[u] = t0 + t1, A[u] = (t0 - t1) x_k pre[v];   phase 2: M[v] = (t0 - t1) eq(r_u; u) x_k, A[v] = (t0 + t1) eq(r_u; u) v_u
template <int PHASE>
__global__ void k_fg_fill_bfly(const F* __restrict__ pre, Pt r0, Pt r1, F alpha, F beta, Pt ru, const F* __restrict__ v_u, const F* __restrict__ tw,
                               int lg, int dep, F* __restrict__ V, F* __restrict__ M, F* __restrict__ A) {
    const uint32_t pairs = 1u << (lg - 1), cols = 1u << dep;
    const F a_lo = f_mul(alpha, f_sub(f_one(), r0.v[lg - 1])), a_hi = f_mul(alpha, r0.v[lg - 1]);
    const F b_lo = f_mul(beta, f_sub(f_one(), r1.v[lg - 1])), b_hi = f_mul(beta, r1.v[lg - 1]);
    const F vu = PHASE == 2 ? ldg(v_u) : f_zero();
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += gridDim.x * blockDim.x) {
        const uint32_t j = q & (cols - 1), k = q >> dep, u = (k << (dep + 1)) | j, v = u | cols, g = (k << dep) | j;
        const F e0 = eq_pt(r0, lg - 1, g), e1 = eq_pt(r1, lg - 1, g);
        const F t0 = f_add(f_mul(a_lo, e0), f_mul(b_lo, e1)), t1 = f_add(f_mul(a_hi, e0), f_mul(b_hi, e1));
        const F x = ldg(tw + (k << dep)), pu = ldg(pre + u), pv = ldg(pre + v);
        stg(V + u, pu);
        stg(V + v, pv);
        if (PHASE == 1) {
            stg(M + u, f_add(t0, t1));
            stg(A + u, f_mul(f_mul(f_sub(t0, t1), x), pv));
            stg(M + v, f_zero());
            stg(A + v, f_zero());
        } else {
            const F eu = eq_pt(ru, lg, u);
            stg(M + v, f_mul(f_mul(f_sub(t0, t1), eu), x));
            stg(A + v, f_mul(f_mul(f_add(t0, t1), eu), vu));
            stg(M + u, f_zero());
            stg(A + u, f_zero());
        }
    }
}

// ---- host side of the field (the verifier's closed forms)
static inline F h_eq(const F* r, int bits, uint64_t g) {
    F a = f_one();
    for (int k = 0; k < bits; ++k) a = f_mul(a, ((g >> k) & 1) ? r[k] : f_sub(f_one(), r[k]));
    return a;
}
static inline F h_pow(F x, unsigned __int128 e) {
    F r = f_one();
    while (e) {
        if (e & 1) r = f_mul(r, x);
        x = f_mul(x, x);
        e >>= 1;
    }
    return r;
}
static inline bool h_same(const F& a, const F& b) { return a.re == b.re && a.im == b.im; }
static inline F h_poly(const F* p, const F& x) { return f_add(f_mul(f_add(f_mul(p[0], x), p[1]), x), p[2]); }
static inline Pt pt_of(const F* r, int n) {
    Pt p;
    for (int i = 0; i < 30; ++i) p.v[i] = i < n ? r[i] : f_zero();
    return p;
}
static inline unsigned grid_of(size_t work) { return (unsigned)std::max<size_t>(1, std::min<size_t>((work + 255) / 256, 148 * 8)); }

static inline size_t rnd_count(int lg) { return (size_t)lg + 64 + 2 * (lg + 10) + 2 * (lg + 6) + 2 * lg + (size_t)lg * (2 * lg + 2); }
static inline size_t poly_count(int lg) { return (size_t)(lg + 6) + lg + (size_t)2 * lg * lg; }

// the 3 n round polynomials + V's final value of the sumcheck that just ran -> the device transcript
static void keep_result(vp_sumcheck* s, F* d_tr, size_t& at) {
    const size_t cnt = (size_t)3 * s->log_n + 1;
    CK(cudaMemcpyAsync(d_tr + at, s->d_out.p, cnt * sizeof(F), cudaMemcpyDeviceToDevice, s->stream));
    at += cnt;
}

using Cache = std::map<std::pair<int, int>, std::pair<vp_sumcheck*, vp_sumcheck*>>;
static std::mutex& cache_mu() { static std::mutex m; return m; }
static Cache& cache() { static Cache c; return c; }
static void release_cache() {
    std::lock_guard<std::mutex> lock(cache_mu());
    for (auto& kv : cache()) {
        if (kv.second.first) vp_sumcheck_destroy(kv.second.first);
        if (kv.second.second) vp_sumcheck_destroy(kv.second.second);
    }
    cache().clear();
}

struct Result {
    int proof_size = 0, ok = 0;
    double verifier_seconds = 0, prover_seconds = 0;
    float device_ms = 0;
};

// rnd: r[lg] | x[64] | r_0[lg+10] | r_1[lg+10] | addition: r_u[lg+6], r_v[lg+6] | mult: r_u[lg], r_v[lg] | per butterfly layer:
// r_u[lg], r_v[lg], alpha, beta. Outputs (host, any may be null): layers (E, F_{lg-1}..F_0, S: (lg+2) 2^lg; P: 64 * 2^lg; O: 64),
// polys (3 per round, protocol order), claims (a_0, after addition, mult, intermediate, each butterfly layer, final alpha, beta).
static Result run(int device, int lg, const F* rnd, F* layers_out, F* polys_out, F* claims_out) {
    Result res;
    const auto t_begin = std::chrono::steady_clock::now();
    CK(cudaSetDevice(device));
    const uint32_t n = 1u << lg;
    // the two sumcheck objects (lg and lg + 6 variables) and their plans are kept for the next call of the same size
    // (release_cache() frees them)
    std::lock_guard<std::mutex> lock(cache_mu());   // (also serialises concurrent callers: the objects hold the working tables)
    auto& slot = cache()[{device, lg}];
    if (!slot.first && sumcheck_create_impl(lg, device, &slot.first, false) != VP_OK) throw CudaError{std::string("fft_gkr: ") + vp_last_error()};
    if (!slot.second && sumcheck_create_impl(lg + 6, device, &slot.second, false) != VP_OK) throw CudaError{std::string("fft_gkr: ") + vp_last_error()};
    vp_sumcheck *s_small = slot.first, *s_big = slot.second;
    cudaStream_t st = s_small->stream;
    cudaEvent_t e0, e1, e_big;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreateWithFlags(&e_big, cudaEventDisableTiming));
    struct EvGuard { cudaEvent_t a, b, c; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); } } evg{e0, e1, e_big};
    // ---- randomness, in draw order
    const F* next = rnd;
    const F* r = next; next += lg;
    const F* xs = next; next += 64;
    std::vector<F> r0(next, next + lg + 10); next += lg + 10;
    std::vector<F> r1(next, next + lg + 10); next += lg + 10;
    // ---- device buffers
    DBuf<F> d_layers, d_P, d_O, d_tw, d_pw, d_small, d_tr;
    d_layers.alloc((size_t)(lg + 2) * n);
    d_P.alloc((size_t)64 * n);
    d_O.alloc(64);
    d_tw.alloc(n);
    d_pw.alloc(64 * 32);
    d_small.alloc(3 * 64);
    const size_t n_polys = poly_count(lg), n_sc = 2 + 2 * (size_t)lg;
    d_tr.alloc(3 * n_polys + n_sc);
    F* E = d_layers.p;
    auto layer = [&](int t) { return d_layers.p + (size_t)t * n; };   // t = 0: E, t = lg - d: F_d, t = lg + 1: S
    CK(cudaEventRecord(e0, st));
    // ---- build_circuit
    k_fg_eq_layer<<<grid_of(n), 256, 0, st>>>(E, lg, pt_of(r, lg));
    F rou{2147483648ULL, 1033321771269002680ULL};                       // order 2^62 (fieldElement.cpp:240-241)
    for (int i = 0; i < 62 - lg; ++i) rou = f_mul(rou, rou);
    const F inv_rou = h_pow(rou, ((unsigned __int128)1 << lg) - 1);     // rou^(2^lg - 1)
    {
        Pt sq;
        F w = inv_rou;
        for (int b = 0; b < 30; ++b) { sq.v[b] = w; w = f_mul(w, w); }
        k_fg_twiddles<<<grid_of(n), 256, 0, st>>>(d_tw.p, lg, sq);
    }
    for (int dep = lg - 1; dep >= 0; --dep)
        k_fg_butterfly<<<grid_of(n / 2), 256, 0, st>>>(layer(lg - dep - 1), layer(lg - dep), lg, dep, d_tw.p);
    const F inv_n = h_pow(F{(u64)n, 0}, (unsigned __int128)P - 2);
    k_fg_scale<<<grid_of(n), 256, 0, st>>>(layer(lg), layer(lg + 1), inv_n, n);
    std::vector<F> pw(64 * 32, f_zero());
    for (int i = 0; i < 64; ++i) {
        F w = xs[i];
        for (int b = 0; b < 32; ++b) { pw[i * 32 + b] = w; w = f_mul(w, w); }
    }
    CK(cudaMemcpyAsync(d_pw.p, pw.data(), pw.size() * sizeof(F), cudaMemcpyHostToDevice, st));
    k_fg_products<<<grid_of((size_t)64 * n), 256, 0, st>>>(layer(lg + 1), d_pw.p, lg, d_P.p);
    k_fg_row_sums<<<64, 256, 0, st>>>(d_P.p, lg, d_O.p);
    // ---- engage_gkr: the prover side of every layer, back to back on the stream
    F alpha = f_one(), beta = f_zero();
    size_t tr_at = 0;
    std::vector<F> small(3 * 64);
    std::vector<std::vector<F>> keep_r0, keep_r1;      // the (r_0, r_1) every layer's verifier check needs
    struct LayerRnd { const F *ru, *rv; F alpha, beta; };
    std::vector<LayerRnd> lr;
    {   // addition layer: lg + 6 rounds on the big object (its stream waits for the circuit, then hands back)
        const F *ru = next; next += lg + 6;
        const F *rv = next; next += lg + 6;
        for (int i = 0; i < 64; ++i) small[i] = f_add(f_mul(alpha, h_eq(r0.data(), 6, i)), f_mul(beta, h_eq(r1.data(), 6, i)));
        CK(cudaMemcpyAsync(d_small.p, small.data(), 64 * sizeof(F), cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(e_big, st));
        CK(cudaStreamWaitEvent(s_big->stream, e_big, 0));
        k_fg_fill_add<<<grid_of((size_t)64 * n), 256, 0, s_big->stream>>>(d_P.p, d_small.p, lg, s_big->bufV[0].p, s_big->bufM[0].p, s_big->bufA[0].p);
        sumcheck_fused_async(s_big, reinterpret_cast<const vp_F*>(ru));
        keep_result(s_big, d_tr.p, tr_at);
        CK(cudaEventRecord(e_big, s_big->stream));
        CK(cudaStreamWaitEvent(st, e_big, 0));
        keep_r0.push_back(r0); keep_r1.push_back(r1);
        lr.push_back({ru, rv, alpha, beta});
        std::copy(ru, ru + lg + 6, r0.begin());
        std::copy(rv, rv + lg + 6, r1.begin());
    }
    {   // mult layer
        const F *ru = next; next += lg;
        const F *rv = next; next += lg;
        for (int j = 0; j < 64; ++j) {
            small[64 + j] = f_mul(alpha, h_eq(r0.data() + lg, 6, j));
            small[128 + j] = f_mul(beta, h_eq(r1.data() + lg, 6, j));
        }
        CK(cudaMemcpyAsync(d_small.p + 64, small.data() + 64, 128 * sizeof(F), cudaMemcpyHostToDevice, st));
        const int use1 = !(beta.re == 0 && beta.im == 0);
        k_fg_fill_mult<<<grid_of(n), 256, 0, st>>>(layer(lg + 1), pt_of(r0.data(), lg), pt_of(r1.data(), lg), d_small.p + 64, d_small.p + 128, use1,
                                                   d_pw.p, lg, s_small->bufV[0].p, s_small->bufM[0].p, s_small->bufA[0].p);
        sumcheck_fused_async(s_small, reinterpret_cast<const vp_F*>(ru));
        keep_result(s_small, d_tr.p, tr_at);
        keep_r0.push_back(r0); keep_r1.push_back(r1);
        lr.push_back({ru, rv, alpha, beta});
        std::copy(ru, ru + lg, r0.begin());
        std::copy(rv, rv + lg, r1.begin());
    }
    for (int dep = 0; dep < lg; ++dep) {   // butterfly layers, output side first
        const F *ru = next; next += lg;
        const F *rv = next; next += lg;
        const F* pre = layer(lg - dep - 1);
        const Pt p0 = pt_of(r0.data(), lg), p1 = pt_of(r1.data(), lg), pu = pt_of(ru, lg);
        k_fg_fill_bfly<1><<<grid_of(n / 2), 256, 0, st>>>(pre, p0, p1, alpha, beta, pu, nullptr, d_tw.p, lg, dep, s_small->bufV[0].p, s_small->bufM[0].p,
                                                           s_small->bufA[0].p);
        sumcheck_fused_async(s_small, reinterpret_cast<const vp_F*>(ru));
        keep_result(s_small, d_tr.p, tr_at);
        const F* v_u = d_tr.p + tr_at - 1;                               // V's final value of phase 1, on the device
        k_fg_fill_bfly<2><<<grid_of(n / 2), 256, 0, st>>>(pre, p0, p1, alpha, beta, pu, v_u, d_tw.p, lg, dep, s_small->bufV[0].p, s_small->bufM[0].p,
                                                           s_small->bufA[0].p);
        sumcheck_fused_async(s_small, reinterpret_cast<const vp_F*>(rv));
        keep_result(s_small, d_tr.p, tr_at);
        keep_r0.push_back(r0); keep_r1.push_back(r1);
        lr.push_back({ru, rv, alpha, beta});
        std::copy(ru, ru + lg, r0.begin());
        std::copy(rv, rv + lg, r1.begin());
        alpha = *next++;
        beta = *next++;
    }
    CK(cudaEventRecord(e1, st));
    CK(cudaGetLastError());
    std::vector<F> tr(d_tr.n), O(64);
    CK(cudaMemcpyAsync(tr.data(), d_tr.p, tr.size() * sizeof(F), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(O.data(), d_O.p, 64 * sizeof(F), cudaMemcpyDeviceToHost, st));
    if (layers_out) {
        CK(cudaMemcpyAsync(layers_out, d_layers.p, d_layers.n * sizeof(F), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(layers_out + d_layers.n, d_P.p, d_P.n * sizeof(F), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(layers_out + d_layers.n + d_P.n, d_O.p, 64 * sizeof(F), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&res.device_ms, e0, e1));
    res.prover_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    // ---- the verifier: claim chain + every layer's wiring predicate in closed form (host)
    const auto tv0 = std::chrono::steady_clock::now();
    bool ok = true;
    int ci = 0, li = 0;
    size_t at = 0, pi = 0;
    F claim;
    {   // V_output :113-130
        std::vector<F> o = O;
        for (int i = 0, sz = 64; i < 6; ++i, sz /= 2)
            for (int j = 0; j < sz / 2; ++j) o[j] = f_add(f_mul(o[2 * j], f_sub(f_one(), keep_r0[0][i])), f_mul(o[2 * j + 1], keep_r0[0][i]));
        claim = o[0];
    }
    auto put_claim = [&](const F& c) { if (claims_out) claims_out[ci] = c; ++ci; };
    put_claim(claim);
    auto rounds = [&](int cnt, const F* rr) {   // p(0) + p(1) == claim, claim = p(r); returns V's final value
        for (int i = 0; i < cnt; ++i) {
            const F* p = tr.data() + at + 3 * (size_t)i;
            if (!h_same(f_add(h_poly(p, f_zero()), h_poly(p, f_one())), claim)) ok = false;
            claim = h_poly(p, rr[i]);
            if (polys_out) std::copy(p, p + 3, polys_out + 3 * pi);
            ++pi;
        }
        at += 3 * (size_t)cnt + 1;
        res.proof_size += 48 * cnt;   // sizeof(quadratic_poly)
        return tr[at - 1];
    };
    {   // addition layer :291-330
        const LayerRnd& Lr = lr[li];
        const std::vector<F>&a0 = keep_r0[li], &a1 = keep_r1[li];
        ++li;
        const F v_u = rounds(lg + 6, Lr.ru);
        F sum = f_zero();
        for (int i = 0; i < 64; ++i)
            sum = f_add(sum, f_mul(f_add(f_mul(Lr.alpha, h_eq(a0.data(), 6, i)), f_mul(Lr.beta, h_eq(a1.data(), 6, i))), h_eq(Lr.ru + lg, 6, i)));
        if (!h_same(claim, f_mul(sum, v_u))) ok = false;
        claim = f_mul(Lr.alpha, v_u);
        put_claim(claim);
    }
    {   // mult layer :404-444
        const LayerRnd& Lr = lr[li];
        const std::vector<F>&a0 = keep_r0[li], &a1 = keep_r1[li];
        ++li;
        const F v_u = rounds(lg, Lr.ru);
        F sum = f_zero();
        for (int i = 0; i < 64; ++i) {
            const F g0 = f_mul(Lr.alpha, h_eq(a0.data() + lg, 6, i)), g1 = f_mul(Lr.beta, h_eq(a1.data() + lg, 6, i));
            F u0 = f_one(), u1 = f_one(), x = xs[i];
            for (int j = 0; j < lg; ++j) {
                const F om_u = f_sub(f_one(), Lr.ru[j]);
                u0 = f_mul(u0, f_add(f_mul(f_mul(a0[j], Lr.ru[j]), x), f_mul(f_sub(f_one(), a0[j]), om_u)));
                u1 = f_mul(u1, f_add(f_mul(f_mul(a1[j], Lr.ru[j]), x), f_mul(f_sub(f_one(), a1[j]), om_u)));
                x = f_mul(x, x);
            }
            sum = f_add(sum, f_add(f_mul(g0, u0), f_mul(g1, u1)));
        }
        if (!h_same(claim, f_mul(sum, v_u))) ok = false;
        claim = f_mul(Lr.alpha, v_u);
        put_claim(claim);
    }
    claim = f_mul(claim, F{(u64)n, 0});   // intermediate layer :447-456
    put_claim(claim);
    F rot = inv_rou;                     // rot_mul[dep] = inv_rou^(2^dep)
    for (int dep = 0; dep < lg; ++dep, rot = f_mul(rot, rot)) {   // :627-765
        const LayerRnd& Lr = lr[li];
        const std::vector<F>&a0 = keep_r0[li], &a1 = keep_r1[li];
        ++li;
        const F v_u = rounds(lg, Lr.ru);
        const F v_v = rounds(lg, Lr.rv);
        // u and v agree with g on the column bits [0, dep) and on the block bits (u, v: (dep, lg); g: [dep, lg-1)); u has bit
        // dep clear, v has it set; g's top bit says upper (+) or lower (-) output of the butterfly
        const F sel = f_mul(f_sub(f_one(), Lr.ru[dep]), Lr.rv[dep]);
        F w_plain[2] = {f_mul(sel, Lr.alpha), f_mul(sel, Lr.beta)}, w_tw[2] = {w_plain[0], w_plain[1]};
        const std::vector<F>* aa[2] = {&a0, &a1};
        F x = rot;
        for (int i = 0; i < lg - dep - 1; ++i, x = f_mul(x, x))
            for (int c = 0; c < 2; ++c) {
                const F g = (*aa[c])[dep + i], u = Lr.ru[dep + 1 + i], v = Lr.rv[dep + 1 + i];
                const F both1 = f_mul(f_mul(g, u), v), both0 = f_mul(f_mul(f_sub(f_one(), g), f_sub(f_one(), u)), f_sub(f_one(), v));
                w_plain[c] = f_mul(w_plain[c], f_add(both1, both0));
                w_tw[c] = f_mul(w_tw[c], f_add(f_mul(both1, x), both0));
            }
        for (int i = 0; i < dep; ++i)
            for (int c = 0; c < 2; ++c) {
                const F g = (*aa[c])[i], u = Lr.ru[i], v = Lr.rv[i];
                const F e = f_add(f_mul(f_mul(g, u), v), f_mul(f_mul(f_sub(f_one(), g), f_sub(f_one(), u)), f_sub(f_one(), v)));
                w_plain[c] = f_mul(w_plain[c], e);
                w_tw[c] = f_mul(w_tw[c], e);
            }
        // upper outputs (top bit of g clear) take +x v, lower ones -x v; both take +u
        F wu = f_zero(), wv = f_zero();
        for (int c = 0; c < 2; ++c) {
            const F top = (*aa[c])[lg - 1];
            wu = f_add(wu, w_plain[c]);                                         // (1 - top) + top
            wv = f_add(wv, f_mul(w_tw[c], f_sub(f_sub(f_one(), top), top)));    // (1 - top) - top
        }
        if (!h_same(claim, f_add(f_mul(wu, v_u), f_mul(wv, v_v)))) ok = false;
        const F na = dep + 1 < lg ? lr[li].alpha : alpha, nb = dep + 1 < lg ? lr[li].beta : beta;
        claim = f_add(f_mul(na, v_u), f_mul(nb, v_v));
        put_claim(claim);
    }
    put_claim(alpha);
    put_claim(beta);
    for (int i = 1; i <= lg; ++i) res.proof_size += 48 * i;   // extension_gkr :771-782
    res.ok = ok ? 1 : 0;
    res.verifier_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tv0).count();
    return res;
}

}  // namespace fg
