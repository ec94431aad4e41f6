// oracle/g16_oracle.cpp -- multithreaded CPU restatement of the Groth16 `prove` hot path of
// microsoft/crescent-credentials (BN254), in the ALGORITHM SHAPE of the arkworks 0.4 crates the reference links.
//
// THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  Nothing under crescent_credentials_b200/ does.
//
// PARITY STATUS: "parity unpinned" against an arkworks binary (no cargo/rustc in the build container, crates not
// vendored).  It is pinned to (a) the in-tree byte goldens (tests/test_oracle_golden.py: zkey.rs:397-431,
// r1cs_reader.rs:266-344), (b) the independent big-integer oracle oracle/pyref.py, result for result.
// It is deliberately a different implementation from the CUDA code: 4 x 64-bit limbs with unsigned __int128 CIOS,
// Jacobian coordinates (ark-ec short_weierstrass::Projective), recursive-free radix-2 FFT.
//
// Reference lines restated:
//   forks/groth16/src/r1cs_to_qap.rs:16-45    evaluate_constraint            -> eval_row
//   forks/groth16/src/r1cs_to_qap.rs:150-213  LibsnarkReduction witness map   -> oc_witness_map (reduction 0)
//   forks/circom-compat/src/circom/qap.rs:25-90  CircomReduction witness map  -> oc_witness_map (reduction 1)
//   forks/groth16/src/prover.rs:54-136,256-274 create_proof_with_assignment    -> oc_prove
//   forks/groth16/src/r1cs_to_qap.rs:106-148  instance_map_with_evaluation    -> oc_instance_map
//   ark-ec 0.4.2 VariableBaseMSM::msm_bigint (signed-digit windows, c = ln(n)+2, one task per window, running-sum
//   bucket reduction, Horner over windows; published algorithm, crate not in tree)   -> msm_pippenger
//   ark-poly 0.4 Radix2EvaluationDomain in-place FFT conventions               -> fft_inplace
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------------------------------
// 256-bit Montgomery fields
// ---------------------------------------------------------------------------------------------------------------------
struct FrTag {
    static constexpr uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    static constexpr uint64_t INV = 0xc2e1f593efffffffull;
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
};
struct FqTag {
    static constexpr uint64_t P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    static constexpr uint64_t INV = 0x87d20782e4866389ull;
    static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
};

template <class T>
struct Fp {
    uint64_t l[4];
    static Fp zero() { return Fp{{0, 0, 0, 0}}; }
    static Fp raw(uint64_t a, uint64_t b = 0, uint64_t c = 0, uint64_t d = 0) { return Fp{{a, b, c, d}}; }
    static Fp r2() { return Fp{{T::R2[0], T::R2[1], T::R2[2], T::R2[3]}}; }
    static Fp one() { return raw(1).to_mont(); }
    static Fp from_u64(uint64_t v) { return raw(v).to_mont(); }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fp& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
    bool operator!=(const Fp& o) const { return !(*this == o); }
    static bool geq_p(const uint64_t* a) {
        for (int i = 3; i >= 0; i--) {
            if (a[i] > T::P[i]) return true;
            if (a[i] < T::P[i]) return false;
        }
        return true;
    }
    static void sub_p(uint64_t* a) {
        u128 b = 0;
        for (int i = 0; i < 4; i++) {
            u128 t = (u128)a[i] - T::P[i] - (uint64_t)b;
            a[i] = (uint64_t)t;
            b = (t >> 64) & 1;
        }
    }
    Fp operator+(const Fp& o) const {
        Fp r;
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)l[i] + o.l[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
        if (c || geq_p(r.l)) sub_p(r.l);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r;
        u128 b = 0;
        for (int i = 0; i < 4; i++) {
            u128 t = (u128)l[i] - o.l[i] - (uint64_t)b;
            r.l[i] = (uint64_t)t;
            b = (t >> 64) & 1;
        }
        if (b) {
            u128 c = 0;
            for (int i = 0; i < 4; i++) {
                c += (u128)r.l[i] + T::P[i];
                r.l[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return r;
    }
    Fp neg() const { return is_zero() ? *this : zero() - *this; }
    Fp dbl() const { return *this + *this; }
    // CIOS Montgomery product on 64-bit limbs
    Fp operator*(const Fp& o) const {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) {
                c += (u128)t[j] + (u128)l[j] * o.l[i];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (uint64_t)c;
            t[5] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * T::INV;
            c = (u128)t[0] + (u128)m * T::P[0];
            c >>= 64;
            for (int j = 1; j < 4; j++) {
                c += (u128)t[j] + (u128)m * T::P[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fp r{{t[0], t[1], t[2], t[3]}};
        if (t[4] || geq_p(r.l)) sub_p(r.l);
        return r;
    }
    Fp sqr() const { return *this * *this; }
    Fp to_mont() const { return *this * r2(); }
    Fp from_mont() const { return *this * raw(1); }
    Fp pow(const uint64_t* e, int limbs) const {
        Fp acc = one();
        for (int i = limbs - 1; i >= 0; i--)
            for (int b = 63; b >= 0; b--) {
                acc = acc.sqr();
                if ((e[i] >> b) & 1) acc = acc * *this;
            }
        return acc;
    }
    Fp pow_u64(uint64_t e) const { return pow(&e, 1); }
    Fp inverse() const {
        uint64_t e[4] = {T::P[0] - 2, T::P[1], T::P[2], T::P[3]};
        return pow(e, 4);
    }
};
typedef Fp<FrTag> Fr;
typedef Fp<FqTag> Fq;

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Fq2& o) const { return !(*this == o); }
    Fq2 operator+(const Fq2& o) const { return Fq2{c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return Fq2{c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return Fq2{c0.neg(), c1.neg()}; }
    Fq2 dbl() const { return Fq2{c0.dbl(), c1.dbl()}; }
    Fq2 operator*(const Fq2& o) const {  // u^2 = -1
        Fq a = c0 * o.c0, b = c1 * o.c1;
        return Fq2{a - b, (c0 + c1) * (o.c0 + o.c1) - a - b};
    }
    Fq2 sqr() const { return *this * *this; }
    Fq2 inverse() const {
        Fq n = (c0.sqr() + c1.sqr()).inverse();
        return Fq2{c0 * n, (c1 * n).neg()};
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// short Weierstrass a = 0, Jacobian coordinates (ark-ec Projective); affine infinity = (0,0)
// ---------------------------------------------------------------------------------------------------------------------
template <class F>
struct Aff {
    F x, y;
    bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
template <class F>
struct Jac {
    F x, y, z;
    static Jac inf() { return Jac{F::one(), F::one(), F::zero()}; }
    bool is_inf() const { return z.is_zero(); }
    static Jac from_affine(const Aff<F>& p) { return p.is_inf() ? inf() : Jac{p.x, p.y, F::one()}; }
    Jac neg() const { return Jac{x, y.neg(), z}; }
    Jac dbl() const {  // dbl-2009-l
        if (is_inf()) return *this;
        F a = x.sqr(), b = y.sqr(), c = b.sqr();
        F d = ((x + b).sqr() - a - c).dbl();
        F e = a.dbl() + a;
        F f = e.sqr();
        Jac r;
        r.z = (y * z).dbl();
        r.x = f - d.dbl();
        r.y = e * (d - r.x) - c.dbl().dbl().dbl();
        return r;
    }
    void add_mixed(const Aff<F>& q) {  // madd-2007-bl
        if (q.is_inf()) return;
        if (is_inf()) {
            *this = from_affine(q);
            return;
        }
        F z1z1 = z.sqr();
        F u2 = q.x * z1z1;
        F s2 = q.y * z * z1z1;
        if (x == u2) {
            if (y == s2) *this = dbl();
            else *this = inf();
            return;
        }
        F h = u2 - x;
        F hh = h.sqr();
        F i = hh.dbl().dbl();
        F j = h * i;
        F r = (s2 - y).dbl();
        F v = x * i;
        F nx = r.sqr() - j - v.dbl();
        F ny = r * (v - nx) - (y * j).dbl();
        F nz = (z + h).sqr() - z1z1 - hh;
        x = nx;
        y = ny;
        z = nz;
    }
    void add(const Jac& q) {  // add-2007-bl
        if (q.is_inf()) return;
        if (is_inf()) {
            *this = q;
            return;
        }
        F z1z1 = z.sqr(), z2z2 = q.z.sqr();
        F u1 = x * z2z2, u2 = q.x * z1z1;
        F s1 = y * q.z * z2z2, s2 = q.y * z * z1z1;
        if (u1 == u2) {
            if (s1 == s2) *this = dbl();
            else *this = inf();
            return;
        }
        F h = u2 - u1;
        F i = h.dbl().sqr();
        F j = h * i;
        F r = (s2 - s1).dbl();
        F v = u1 * i;
        F nx = r.sqr() - j - v.dbl();
        F ny = r * (v - nx) - (s1 * j).dbl();
        F nz = ((z + q.z).sqr() - z1z1 - z2z2) * h;
        x = nx;
        y = ny;
        z = nz;
    }
    Aff<F> to_affine() const {
        if (is_inf()) return Aff<F>{F::zero(), F::zero()};
        F zi = z.inverse();
        F zi2 = zi.sqr();
        return Aff<F>{x * zi2, y * zi2 * zi};
    }
    // mul_bigint: plain double-and-add over a 256-bit canonical scalar
    Jac mul(const uint64_t* k) const {
        Jac acc = inf();
        for (int i = 3; i >= 0; i--)
            for (int b = 63; b >= 0; b--) {
                acc = acc.dbl();
                if ((k[i] >> b) & 1) acc.add(*this);
            }
        return acc;
    }
};
typedef Aff<Fq> G1A;
typedef Aff<Fq2> G2A;
typedef Jac<Fq> G1J;
typedef Jac<Fq2> G2J;

static G1A g1_gen() { return G1A{Fq::from_u64(1), Fq::from_u64(2)}; }
static G2A g2_gen() {
    // forks/circom-compat/src/zkey.rs:442-462
    Fq x0 = Fq::raw(0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull).to_mont();
    Fq x1 = Fq::raw(0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull).to_mont();
    Fq y0 = Fq::raw(0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull).to_mont();
    Fq y1 = Fq::raw(0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull).to_mont();
    return G2A{Fq2{x0, x1}, Fq2{y0, y1}};
}

// ---------------------------------------------------------------------------------------------------------------------
// threading helper
// ---------------------------------------------------------------------------------------------------------------------
static void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t)>& fn) {
    if (threads <= 1 || n < 2) {
        fn(0, n);
        return;
    }
    size_t nt = std::min<size_t>((size_t)threads, n);
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) {
        size_t lo = n * t / nt, hi = n * (t + 1) / nt;
        th.emplace_back([=, &fn] { fn(lo, hi); });
    }
    for (auto& t : th) t.join();
}
// dynamic scheduling of `n` independent tasks over `threads` workers (rayon-like work distribution)
static void parallel_tasks(size_t n, int threads, const std::function<void(size_t)>& fn) {
    std::atomic<size_t> next(0);
    size_t nt = std::max<size_t>(1, std::min<size_t>((size_t)threads, n));
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&] {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                fn(i);
            }
        });
    for (auto& t : th) t.join();
}

// ---------------------------------------------------------------------------------------------------------------------
// domain + FFT (ark-poly conventions): omega = rho^(2^(28 - log n)), rho = 5^((r-1)/2^28); g = 5
// ---------------------------------------------------------------------------------------------------------------------
static Fr fr_generator() { return Fr::from_u64(5); }
static Fr fr_root_of_unity(unsigned log_n) {
    // (r - 1) / 2^28
    static const uint64_t e[4] = {0x9b9709143e1f593full, 0x181585d2833e8487ull, 0x131a029b85045b68ull, 0x000000030644e72eull};
    Fr w = fr_generator().pow(e, 4);
    for (unsigned i = log_n; i < 28; i++) w = w.sqr();
    return w;
}

static void bit_reverse(Fr* a, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) {
        size_t r = 0;
        for (unsigned b = 0; b < log_n; b++)
            if (i >> b & 1) r |= (size_t)1 << (log_n - 1 - b);
        if (i < r) std::swap(a[i], a[r]);
    }
}

// in-place radix-2 DIT, natural order in/out; `w` generates the size-n subgroup
static void fft_inplace(Fr* a, unsigned log_n, Fr w, int threads) {
    size_t n = (size_t)1 << log_n;
    if (n == 1) return;
    bit_reverse(a, log_n);
    // twiddle table w^k, k < n/2
    std::vector<Fr> tw(n / 2);
    {
        size_t chunk = 1 << 12;
        size_t nch = (n / 2 + chunk - 1) / chunk;
        parallel_tasks(nch, threads, [&](size_t c) {
            size_t lo = c * chunk, hi = std::min(n / 2, lo + chunk);
            Fr cur = w.pow_u64(lo);
            for (size_t k = lo; k < hi; k++) {
                tw[k] = cur;
                cur = cur * w;
            }
        });
    }
    for (unsigned s = 0; s < log_n; s++) {
        size_t half = (size_t)1 << s;
        size_t stride = (n / 2) >> s;
        parallel_for(n / 2, threads, [&](size_t lo, size_t hi) {
            for (size_t b = lo; b < hi; b++) {
                size_t k = b & (half - 1);
                size_t i0 = ((b >> s) << (s + 1)) | k;
                size_t i1 = i0 + half;
                Fr t = a[i1] * tw[k * stride];
                Fr u = a[i0];
                a[i0] = u + t;
                a[i1] = u - t;
            }
        });
    }
}

struct Domain {
    unsigned log_n;
    size_t n;
    Fr w, w_inv, n_inv;
    explicit Domain(unsigned lg) : log_n(lg), n((size_t)1 << lg) {
        w = fr_root_of_unity(lg);
        w_inv = w.inverse();
        n_inv = Fr::from_u64(n).inverse();
    }
    void fft(Fr* a, int th) const { fft_inplace(a, log_n, w, th); }
    void ifft(Fr* a, int th) const {
        fft_inplace(a, log_n, w_inv, th);
        parallel_for(n, th, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) a[i] = a[i] * n_inv;
        });
    }
    static void distribute_powers(Fr* a, size_t n, Fr g, Fr c, int th) {  // a[i] *= c * g^i
        parallel_for(n, th, [&](size_t lo, size_t hi) {
            Fr p = c * g.pow_u64(lo);
            for (size_t i = lo; i < hi; i++) {
                a[i] = a[i] * p;
                p = p * g;
            }
        });
    }
    void coset_fft(Fr* a, Fr g, int th) const {
        distribute_powers(a, n, g, Fr::one(), th);
        fft(a, th);
    }
    void coset_ifft(Fr* a, Fr g, int th) const {
        ifft(a, th);
        distribute_powers(a, n, g.inverse(), Fr::one(), th);
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// R1CS
// ---------------------------------------------------------------------------------------------------------------------
struct Csr {
    const uint64_t* row_ptr;
    const uint32_t* col;
    const Fr* val;
};

static Fr eval_row(const Csr& m, size_t i, const Fr* z, const Fr& one) {
    Fr s = Fr::zero();
    for (uint64_t k = m.row_ptr[i]; k < m.row_ptr[i + 1]; k++) {
        const Fr& c = m.val[k];
        const Fr& v = z[m.col[k]];
        if (c == one) s = s + v;
        else s = s + v * c;
    }
    return s;
}

static unsigned domain_log(uint64_t need) {
    unsigned lg = 0;
    while (((uint64_t)1 << lg) < need) lg++;
    return lg;
}

static int witness_map(const Csr m[3], uint64_t nc, uint64_t ni, const Fr* z, int reduction, Fr* h, int th) {
    unsigned lg = domain_log(nc + ni);
    if (lg > 28) return 1;
    Domain dom(lg);
    size_t n = dom.n;
    Fr one = Fr::one();
    std::vector<Fr> a(n, Fr::zero()), b(n, Fr::zero()), c(n, Fr::zero());
    parallel_for(nc, th, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            a[i] = eval_row(m[0], i, z, one);
            b[i] = eval_row(m[1], i, z, one);
        }
    });
    for (uint64_t i = 0; i < ni; i++) a[nc + i] = z[i];
    if (reduction == 0) {
        dom.ifft(a.data(), th);
        dom.ifft(b.data(), th);
        Fr g = fr_generator();
        dom.coset_fft(a.data(), g, th);
        dom.coset_fft(b.data(), g, th);
        parallel_for(n, th, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) a[i] = a[i] * b[i];
        });
        parallel_for(nc, th, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) c[i] = eval_row(m[2], i, z, one);
        });
        dom.ifft(c.data(), th);
        dom.coset_fft(c.data(), g, th);
        Fr zv = g.pow_u64(n) - one;
        if (zv.is_zero()) return 6;
        Fr zi = zv.inverse();
        parallel_for(n, th, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) a[i] = (a[i] - c[i]) * zi;
        });
        dom.coset_ifft(a.data(), g, th);
    } else {
        if (lg >= 28) return 1;
        parallel_for(nc, th, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) c[i] = a[i] * b[i];
        });
        dom.ifft(a.data(), th);
        dom.ifft(b.data(), th);
        Fr root = fr_root_of_unity(lg + 1);
        Domain::distribute_powers(a.data(), n, root, one, th);
        Domain::distribute_powers(b.data(), n, root, one, th);
        dom.fft(a.data(), th);
        dom.fft(b.data(), th);
        parallel_for(n, th, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) a[i] = a[i] * b[i];
        });
        dom.ifft(c.data(), th);
        Domain::distribute_powers(c.data(), n, root, one, th);
        dom.fft(c.data(), th);
        parallel_for(n, th, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) a[i] = a[i] - c[i];
        });
    }
    memcpy(h, a.data(), n * sizeof(Fr));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Pippenger in arkworks' shape
// ---------------------------------------------------------------------------------------------------------------------
static unsigned ark_window(size_t n) {
    if (n < 32) return 3;
    unsigned lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    // ln_without_floats: log2(n) * 69 / 100, with ark's log2 = ceil-ish (64 - leading_zeros(n)) ... use their formula
    unsigned log2c = 0;
    {
        size_t a = n;
        while (a) {
            log2c++;
            a >>= 1;
        }
    }
    (void)lg;
    return (log2c * 69 / 100) + 2;
}

template <class F>
static Jac<F> msm_pippenger(const Aff<F>* bases, const Fr* scalars_mont, size_t n, int threads) {
    // canonical scalars, zero scalars filtered (ark-ec msm_bigint)
    std::vector<std::array<uint64_t, 4>> sc;
    std::vector<const Aff<F>*> pts;
    sc.reserve(n);
    pts.reserve(n);
    for (size_t i = 0; i < n; i++) {
        Fr c = scalars_mont[i].from_mont();
        if (c.is_zero()) continue;
        sc.push_back({c.l[0], c.l[1], c.l[2], c.l[3]});
        pts.push_back(&bases[i]);
    }
    size_t m = sc.size();
    if (m == 0) return Jac<F>::inf();
    unsigned c = ark_window(m);
    const unsigned num_bits = 254;
    unsigned digits_count = (num_bits + c - 1) / c;
    // signed radix-2^c digits (ark-ec make_digits)
    std::vector<int32_t> digits((size_t)m * digits_count);
    parallel_for(m, threads, [&](size_t lo, size_t hi) {
        const uint64_t radix = (uint64_t)1 << c;
        const uint64_t window_mask = radix - 1;
        for (size_t i = lo; i < hi; i++) {
            uint64_t carry = 0;
            for (unsigned d = 0; d < digits_count; d++) {
                unsigned bit_offset = d * c;
                unsigned u64_idx = bit_offset / 64, bit_idx = bit_offset % 64;
                uint64_t bit_buf;
                if (bit_idx < 64 - c || u64_idx == 3)
                    bit_buf = sc[i][u64_idx] >> bit_idx;
                else
                    bit_buf = (sc[i][u64_idx] >> bit_idx) | (sc[i][u64_idx + 1] << (64 - bit_idx));
                uint64_t coef = carry + (bit_buf & window_mask);
                carry = (coef + radix / 2) >> c;
                int64_t dig = (int64_t)coef - (int64_t)(carry << c);
                digits[i * digits_count + d] = (int32_t)dig;
            }
            // ark adds the final carry into the last digit
            digits[i * digits_count + digits_count - 1] += (int32_t)(carry << c);
        }
    });
    std::vector<Jac<F>> window_sums(digits_count, Jac<F>::inf());
    parallel_tasks(digits_count, threads, [&](size_t w) {
        std::vector<Jac<F>> buckets((size_t)1 << (c - 1), Jac<F>::inf());
        for (size_t i = 0; i < m; i++) {
            int32_t d = digits[i * digits_count + w];
            if (d > 0) buckets[d - 1].add_mixed(*pts[i]);
            else if (d < 0) {
                Aff<F> np{pts[i]->x, pts[i]->y.neg()};
                if (!pts[i]->is_inf()) buckets[-d - 1].add_mixed(np);
            }
        }
        Jac<F> run = Jac<F>::inf(), res = Jac<F>::inf();
        for (size_t b = buckets.size(); b-- > 0;) {
            run.add(buckets[b]);
            res.add(run);
        }
        window_sums[w] = res;
    });
    Jac<F> total = Jac<F>::inf();
    for (unsigned w = digits_count; w-- > 1;) {
        total.add(window_sums[w]);
        for (unsigned k = 0; k < c; k++) total = total.dbl();
    }
    total.add(window_sums[0]);
    return total;
}

// naive double-and-add reference (independent of the bucket method) for small n
template <class F>
static Jac<F> msm_naive(const Aff<F>* bases, const Fr* scalars_mont, size_t n) {
    Jac<F> acc = Jac<F>::inf();
    for (size_t i = 0; i < n; i++) {
        Fr c = scalars_mont[i].from_mont();
        acc.add(Jac<F>::from_affine(bases[i]).mul(c.l));
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// C interface (ctypes)
// ---------------------------------------------------------------------------------------------------------------------
struct oc_pk {
    const G1A* a_query; size_t a_len;
    const G1A* b_g1_query; size_t b_g1_len;
    const G2A* b_g2_query; size_t b_g2_len;
    const G1A* h_query; size_t h_len;
    const G1A* l_query; size_t l_len;
    const G1A* alpha_g1; const G1A* beta_g1; const G1A* delta_g1;
    const G2A* beta_g2; const G2A* delta_g2;
};
struct oc_r1cs {
    uint64_t nc, ni, m;
    const uint64_t* row_ptr[3];
    const uint32_t* col[3];
    const Fr* val[3];
};
struct oc_proof {
    G1A a;
    G2A b;
    G1A c;
};
struct oc_timings {
    double witness_map_s, msm_h_s, msm_l_s, msm_a_s, msm_b_g1_s, msm_b_g2_s, total_s;
};

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class F>
static Jac<F> calculate_coeff(const Jac<F>& initial, const Aff<F>* query, size_t qlen, const Aff<F>& vk_param, const Fr* assignment,
                              size_t alen, int th) {
    // prover.rs:256-274
    size_t n = std::min(qlen - 1, alen);
    Jac<F> acc = msm_pippenger<F>(query + 1, assignment, n, th);
    Jac<F> res = initial;
    res.add_mixed(query[0]);
    res.add(acc);
    res.add_mixed(vk_param);
    return res;
}

extern "C" {

int oc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
template <class F>
static void field_op_range(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t lo, size_t hi) {
    const bool bcast = op == 8 || op == 9;  // b is ONE element (same op codes as include/g16_b200.h: G16_OP_*)
    F yb = F::zero();
    if (bcast && b) memcpy(&yb, b, 32);
    for (size_t i = lo; i < hi; i++) {
        F x, y = yb, z;
        memcpy(&x, a + 4 * i, 32);
        if (b && !bcast) memcpy(&y, b + 4 * i, 32);
        switch (op) {
            case 0: case 8: z = x * y; break;
            case 1: case 9: z = x + y; break;
            case 2: z = x - y; break;
            case 3: z = x.neg(); break;
            case 4: z = x.inverse(); break;
            case 5: z = x.to_mont(); break;
            case 6: z = x.from_mont(); break;
            default: z = x.sqr();
        }
        memcpy(out + 4 * i, &z, 32);
    }
}

extern "C" {
void oc_field_op_mt(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, int threads) {
    parallel_for(n, threads, [&](size_t lo, size_t hi) {
        if (field == 0) field_op_range<Fr>(op, a, b, out, lo, hi);
        else field_op_range<Fq>(op, a, b, out, lo, hi);
    });
}
void oc_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    oc_field_op_mt(field, op, a, b, out, n, n >= 4096 ? (int)std::thread::hardware_concurrency() : 1);
}

// out[i] = scale * base^i (Montgomery in / out)
void oc_pow_table(const uint64_t* base, const uint64_t* scale, size_t n, uint64_t* out, int threads) {
    Fr b, s;
    memcpy(&b, base, 32);
    memcpy(&s, scale, 32);
    Fr* o = (Fr*)out;
    parallel_for(n, threads, [&](size_t lo, size_t hi) {
        Fr cur = s * b.pow_u64(lo);
        for (size_t i = lo; i < hi; i++) {
            o[i] = cur;
            cur = cur * b;
        }
    });
}

// arkworks-semantics transform, natural order in/out, in place
int oc_ntt(uint64_t* data, unsigned log_n, int inverse, int coset, int threads) {
    if (log_n > 28) return 1;
    Domain d(log_n);
    Fr* a = (Fr*)data;
    Fr g = fr_generator();
    if (!inverse) {
        if (coset) d.coset_fft(a, g, threads);
        else d.fft(a, threads);
    } else {
        if (coset) d.coset_ifft(a, g, threads);
        else d.ifft(a, threads);
    }
    return 0;
}

int oc_witness_map(const oc_r1cs* r, const uint64_t* z, int reduction, uint64_t* h, int threads) {
    Csr m[3];
    for (int k = 0; k < 3; k++) m[k] = Csr{r->row_ptr[k], r->col[k], r->val[k]};
    return witness_map(m, r->nc, r->ni, (const Fr*)z, reduction, (Fr*)h, threads);
}

void oc_r1cs_eval(const oc_r1cs* r, const uint64_t* z, uint64_t* az, uint64_t* bz, uint64_t* cz, int threads) {
    Fr one = Fr::one();
    uint64_t* outs[3] = {az, bz, cz};
    for (int k = 0; k < 3; k++) {
        if (!outs[k]) continue;
        Csr m{r->row_ptr[k], r->col[k], r->val[k]};
        Fr* o = (Fr*)outs[k];
        parallel_for(r->nc, threads, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) o[i] = eval_row(m, i, (const Fr*)z, one);
        });
    }
}

void oc_msm_g1(const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t* out, int naive, int threads) {
    G1J r = naive ? msm_naive<Fq>((const G1A*)points, (const Fr*)scalars, n)
                  : msm_pippenger<Fq>((const G1A*)points, (const Fr*)scalars, n, threads);
    G1A a = r.to_affine();
    memcpy(out, &a, sizeof(a));
}
void oc_msm_g2(const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t* out, int naive, int threads) {
    G2J r = naive ? msm_naive<Fq2>((const G2A*)points, (const Fr*)scalars, n)
                  : msm_pippenger<Fq2>((const G2A*)points, (const Fr*)scalars, n, threads);
    G2A a = r.to_affine();
    memcpy(out, &a, sizeof(a));
}

}  // extern "C"
// out[i] = k_i * G (canonical generators), via an 8-bit fixed-base window table; Jacobian results are normalised
// 256 at a time with one shared inversion (ark-ec's batch normalisation, Projective::normalize_batch)
template <class F>
static void fixed_base_t(const Aff<F>& gen, const Fr* scalars, size_t n, Aff<F>* out, int threads) {
    const unsigned W = 8, NW = 32;
    std::vector<Aff<F>> tbl((size_t)NW << W);
    Jac<F> base = Jac<F>::from_affine(gen);
    for (unsigned w = 0; w < NW; w++) {
        Jac<F> cur = Jac<F>::inf();
        tbl[(size_t)w << W] = Aff<F>{F::zero(), F::zero()};
        for (unsigned d = 1; d < (1u << W); d++) {
            cur.add(base);
            tbl[((size_t)w << W) + d] = cur.to_affine();
        }
        for (unsigned k = 0; k < W; k++) base = base.dbl();
    }
    const size_t B = 256;
    parallel_tasks((n + B - 1) / B, threads, [&](size_t blk) {
        size_t lo = blk * B, hi = std::min(n, lo + B);
        Jac<F> acc[B];
        F pre[B];
        F run = F::one();
        for (size_t i = lo; i < hi; i++) {
            Fr c = scalars[i].from_mont();
            Jac<F> a = Jac<F>::inf();
            for (unsigned w = 0; w < NW; w++) {
                unsigned d = (c.l[w / 8] >> ((w % 8) * 8)) & 0xff;
                if (d) a.add_mixed(tbl[((size_t)w << W) + d]);
            }
            acc[i - lo] = a;
            pre[i - lo] = run;
            if (!a.is_inf()) run = run * a.z;
        }
        F inv = run.inverse();
        for (size_t i = hi; i-- > lo;) {
            const Jac<F>& a = acc[i - lo];
            if (a.is_inf()) {
                out[i] = Aff<F>{F::zero(), F::zero()};
                continue;
            }
            F zi = inv * pre[i - lo];
            inv = inv * a.z;
            F zi2 = zi.sqr();
            out[i] = Aff<F>{a.x * zi2, a.y * zi2 * zi};
        }
    });
}
extern "C" {
void oc_fixed_base(int group, const uint64_t* scalars, size_t n, uint64_t* out, int threads) {
    if (group == 1) fixed_base_t<Fq>(g1_gen(), (const Fr*)scalars, n, (G1A*)out, threads);
    else fixed_base_t<Fq2>(g2_gen(), (const Fr*)scalars, n, (G2A*)out, threads);
}

// instance_map_with_evaluation (r1cs_to_qap.rs:106-148): a, b, c receive m Montgomery elements each; returns Z(t) in zt
int oc_instance_map(const oc_r1cs* r, const uint64_t* t_mont, uint64_t* a_out, uint64_t* b_out, uint64_t* c_out, uint64_t* zt_out,
                    uint64_t* domain_size_out, int threads) {
    unsigned lg = domain_log(r->nc + r->ni);
    if (lg > 28) return 1;
    Domain dom(lg);
    size_t n = dom.n;
    Fr t;
    memcpy(&t, t_mont, 32);
    Fr one = Fr::one();
    Fr zt = t.pow_u64(n) - one;
    // L_i(t) = Z(t) * w^i / (n * (t - w^i)); batch inversion per chunk
    std::vector<Fr> u(n);
    size_t chunk = 1 << 12;
    parallel_tasks((n + chunk - 1) / chunk, threads, [&](size_t cix) {
        size_t lo = cix * chunk, hi = std::min(n, lo + chunk);
        std::vector<Fr> den(hi - lo), pre(hi - lo), wi(hi - lo);
        Fr w = dom.w.pow_u64(lo);
        Fr nf = Fr::from_u64(n);
        Fr acc = one;
        for (size_t i = lo; i < hi; i++) {
            wi[i - lo] = w;
            den[i - lo] = nf * (t - w);
            pre[i - lo] = acc;
            acc = acc * den[i - lo];
            w = w * dom.w;
        }
        Fr inv = acc.inverse();
        for (size_t i = hi; i-- > lo;) {
            Fr di = inv * pre[i - lo];
            inv = inv * den[i - lo];
            u[i] = zt * wi[i - lo] * di;
        }
    });
    Fr* a = (Fr*)a_out;
    Fr* b = (Fr*)b_out;
    Fr* c = (Fr*)c_out;
    for (uint64_t i = 0; i < r->m; i++) a[i] = b[i] = c[i] = Fr::zero();
    for (uint64_t i = 0; i < r->ni; i++) a[i] = u[r->nc + i];
    Fr* outs[3] = {a, b, c};
    // column scatter: three matrices in parallel (each owns its output vector)
    parallel_tasks(3, threads, [&](size_t k) {
        Fr* o = outs[k];
        for (uint64_t i = 0; i < r->nc; i++)
            for (uint64_t p = r->row_ptr[k][i]; p < r->row_ptr[k][i + 1]; p++) {
                const Fr& cf = r->val[k][p];
                o[r->col[k][p]] = o[r->col[k][p]] + (cf == one ? u[i] : u[i] * cf);
            }
    });
    memcpy(zt_out, &zt, 32);
    *domain_size_out = n;
    return 0;
}

// Groth16::create_proof_with_reduction_and_matrices (prover.rs:26-51) on the CPU
int oc_prove(const oc_pk* pk, const oc_r1cs* r, const uint64_t* z_mont, const uint64_t* r_mont, const uint64_t* s_mont, int reduction,
             oc_proof* out, uint64_t* h_out /* optional, n elements */, oc_timings* tm, int threads) {
    double t0 = now_s();
    unsigned lg = domain_log(r->nc + r->ni);
    if (lg > 28) return 1;
    size_t n = (size_t)1 << lg;
    const Fr* z = (const Fr*)z_mont;
    std::vector<Fr> h(n);
    Csr m[3];
    for (int k = 0; k < 3; k++) m[k] = Csr{r->row_ptr[k], r->col[k], r->val[k]};
    int rc = witness_map(m, r->nc, r->ni, z, reduction, h.data(), threads);
    if (rc) return rc;
    double t1 = now_s();
    if (h_out) memcpy(h_out, h.data(), n * 32);
    Fr rr, ss;
    memcpy(&rr, r_mont, 32);
    memcpy(&ss, s_mont, 32);
    Fr rc_ = rr.from_mont(), sc_ = ss.from_mont();
    // prover.rs:63-66, 70-74
    G1J h_acc = msm_pippenger<Fq>(pk->h_query, h.data(), std::min(pk->h_len, n), threads);
    double t2 = now_s();
    const Fr* aux = z + r->ni;
    size_t naux = r->m - r->ni;
    G1J l_acc = msm_pippenger<Fq>(pk->l_query, aux, std::min(pk->l_len, naux), threads);
    double t3 = now_s();
    G1J delta1 = G1J::from_affine(*pk->delta_g1);
    G1J rs_delta = delta1.mul(rc_.l).mul(sc_.l);  // :76-80
    const Fr* assignment = z + 1;                 // input_assignment ++ aux_assignment (:84-89)
    size_t alen = r->m - 1;
    G1J r_g1 = delta1.mul(rc_.l);  // :94
    G1J g_a = calculate_coeff<Fq>(r_g1, pk->a_query, pk->a_len, *pk->alpha_g1, assignment, alen, threads);
    G1J s_g_a = g_a.mul(sc_.l);
    double t4 = now_s();
    G1J g1_b = G1J::inf();
    if (!rr.is_zero()) {  // :102-112
        G1J s_g1 = delta1.mul(sc_.l);
        g1_b = calculate_coeff<Fq>(s_g1, pk->b_g1_query, pk->b_g1_len, *pk->beta_g1, assignment, alen, threads);
    }
    double t5 = now_s();
    G2J s_g2 = G2J::from_affine(*pk->delta_g2).mul(sc_.l);  // :116
    G2J g2_b = calculate_coeff<Fq2>(s_g2, pk->b_g2_query, pk->b_g2_len, *pk->beta_g2, assignment, alen, threads);
    G1J r_g1_b = g1_b.mul(rc_.l);
    double t6 = now_s();
    G1J g_c = s_g_a;  // :124-128
    g_c.add(r_g1_b);
    g_c.add(rs_delta.neg());
    g_c.add(l_acc);
    g_c.add(h_acc);
    out->a = g_a.to_affine();
    out->b = g2_b.to_affine();
    out->c = g_c.to_affine();
    double t7 = now_s();
    if (tm) {
        tm->witness_map_s = t1 - t0;
        tm->msm_h_s = t2 - t1;
        tm->msm_l_s = t3 - t2;
        tm->msm_a_s = t4 - t3;
        tm->msm_b_g1_s = t5 - t4;
        tm->msm_b_g2_s = t6 - t5;
        tm->total_s = t7 - t0;
    }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// The verifier (forks/groth16/src/verifier.rs) and the BN254 pairing it calls, in the shape ark-ec 0.4 `models::bn` gives
// them on the CPU: G2Prepared line coefficients for all three pairs, multi_miller_loop over the signed digits of 6x+2,
// final exponentiation with cyclotomic squarings and the Fuentes-Castaneda hard part; prepare_inputs as one plain
// double-and-add `mul_bigint` per public input (verifier.rs:33-36).  Third implementation of this arithmetic next to
// oracle/pairing.py (big integers, Fq12 = Fq2[w]/(w^6 - xi)) and csrc/pairing.cuh (32-bit limbs): 64-bit limbs, tower
// Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v) with every constant derived at start-up from q and xi = 9 + u.
// Used as a checker and as the CPU baseline of tools/verify_bench.py.
// ---------------------------------------------------------------------------------------------------------------------
namespace pairing_oracle {

static Fq2 mul_xi(const Fq2& a) {  // (9 + u) a
    Fq t0 = a.c0.dbl().dbl().dbl() + a.c0, t1 = a.c1.dbl().dbl().dbl() + a.c1;
    return Fq2{t0 - a.c1, t1 + a.c0};
}
static Fq2 conj2(const Fq2& a) { return Fq2{a.c0, a.c1.neg()}; }
static Fq2 scale2(const Fq2& a, const Fq& k) { return Fq2{a.c0 * k, a.c1 * k}; }

struct Fq6 {
    Fq2 c0, c1, c2;
    static Fq6 zero() { return Fq6{Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static Fq6 one() { return Fq6{Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
    bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    Fq6 operator+(const Fq6& o) const { return Fq6{c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fq6 operator-(const Fq6& o) const { return Fq6{c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fq6 neg() const { return Fq6{c0.neg(), c1.neg(), c2.neg()}; }
    Fq6 mul_v() const { return Fq6{mul_xi(c2), c0, c1}; }
    Fq6 operator*(const Fq6& o) const {  // schoolbook: 9 Fq2 products (the device code uses Karatsuba)
        Fq2 r0 = c0 * o.c0 + mul_xi(c1 * o.c2 + c2 * o.c1);
        Fq2 r1 = c0 * o.c1 + c1 * o.c0 + mul_xi(c2 * o.c2);
        Fq2 r2 = c0 * o.c2 + c1 * o.c1 + c2 * o.c0;
        return Fq6{r0, r1, r2};
    }
    Fq6 inverse() const {
        Fq2 t0 = c0.sqr() - mul_xi(c1 * c2), t1 = mul_xi(c2.sqr()) - c0 * c1, t2 = c1.sqr() - c0 * c2;
        Fq2 d = (c0 * t0 + mul_xi(c2 * t1 + c1 * t2)).inverse();
        return Fq6{t0 * d, t1 * d, t2 * d};
    }
};

struct Fq12 {
    Fq6 c0, c1;
    static Fq12 one() { return Fq12{Fq6::one(), Fq6::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq12 operator*(const Fq12& o) const {
        Fq6 a = c0 * o.c0, b = c1 * o.c1;
        return Fq12{a + b.mul_v(), (c0 + c1) * (o.c0 + o.c1) - a - b};
    }
    Fq12 sqr() const { return *this * *this; }
    Fq12 conj() const { return Fq12{c0, c1.neg()}; }
    Fq12 inverse() const {
        Fq6 n = (c0 * c0 - (c1 * c1).mul_v()).inverse();
        return Fq12{c0 * n, (c1 * n).neg()};
    }
};

struct Consts {
    Fq2 frob[3][6];  // frob[n-1][k] = xi^(k (q^n - 1)/6)
    Fq2 twist_x, twist_y, b2;
    Fq two_inv;
    Consts() {
        // (q - 1) / 6 by long division of the four limbs
        uint64_t e[4];
        u128 rem = 0;
        uint64_t qm1[4] = {FqTag::P[0] - 1, FqTag::P[1], FqTag::P[2], FqTag::P[3]};
        for (int i = 3; i >= 0; i--) {
            u128 cur = (rem << 64) | qm1[i];
            e[i] = (uint64_t)(cur / 6);
            rem = cur % 6;
        }
        Fq2 xi{Fq::from_u64(9), Fq::one()};
        Fq2 c1 = Fq2::one();
        for (int i = 3; i >= 0; i--)
            for (int b = 63; b >= 0; b--) {
                c1 = c1.sqr();
                if ((e[i] >> b) & 1) c1 = c1 * xi;
            }
        frob[0][0] = frob[1][0] = frob[2][0] = Fq2::one();
        for (int k = 1; k < 6; k++) frob[0][k] = frob[0][k - 1] * c1;
        for (int k = 1; k < 6; k++) {
            frob[1][k] = frob[0][k] * conj2(frob[0][k]);  // f^(q + 1)
            frob[2][k] = frob[1][k] * frob[0][k];         // f^(q^2 + q + 1), f^(q^2) = f
        }
        twist_x = frob[0][2];
        twist_y = frob[0][3];
        b2 = scale2(xi.inverse(), Fq::from_u64(3));
        two_inv = Fq::from_u64(2).inverse();
    }
};
static const Consts& K() {
    static const Consts k;
    return k;
}

static Fq12 frobenius(const Fq12& a, int n) {
    const Fq2* f = K().frob[n - 1];
    auto m = [&](const Fq2& x, int k) { return ((n & 1) ? conj2(x) : x) * f[k]; };
    return Fq12{Fq6{m(a.c0.c0, 0), m(a.c0.c1, 2), m(a.c0.c2, 4)}, Fq6{m(a.c1.c0, 1), m(a.c1.c1, 3), m(a.c1.c2, 5)}};
}

// Granger-Scott squaring in the cyclotomic subgroup (ark-ff Fp12::cyclotomic_square)
static Fq12 cyclotomic_sqr(const Fq12& a) {
    auto sq4 = [](const Fq2& x, const Fq2& y, Fq2& r0, Fq2& r1) {
        Fq2 t = x * y;
        r0 = (x + y) * (mul_xi(y) + x) - t - mul_xi(t);
        r1 = t.dbl();
    };
    Fq2 t0, t1, t2, t3, t4, t5;
    sq4(a.c0.c0, a.c1.c1, t0, t1);
    sq4(a.c1.c0, a.c0.c2, t2, t3);
    sq4(a.c0.c1, a.c1.c2, t4, t5);
    auto s = [](const Fq2& t, const Fq2& z) { return (t - z).dbl() + t; };
    auto p = [](const Fq2& t, const Fq2& z) { return (t + z).dbl() + t; };
    Fq12 r;
    r.c0.c0 = s(t0, a.c0.c0);
    r.c1.c1 = p(t1, a.c1.c1);
    r.c1.c0 = p(mul_xi(t5), a.c1.c0);
    r.c0.c2 = s(t4, a.c0.c2);
    r.c0.c1 = s(t2, a.c0.c1);
    r.c1.c2 = p(t3, a.c1.c2);
    return r;
}

static const int8_t ATE[65] = {0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0, 1, 1, 1,
                               0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1};
static const uint64_t BN_X = 4965661367192848881ull;

struct Ell {
    Fq2 c0, c1, c2;
};
struct Hom {
    Fq2 x, y, z;
};
static Ell double_step(Hom& r) {
    const Consts& k = K();
    Fq2 a = scale2(r.x * r.y, k.two_inv), b = r.y.sqr(), c = r.z.sqr();
    Fq2 e = k.b2 * (c.dbl() + c), f = e.dbl() + e;
    Fq2 g = scale2(b + f, k.two_inv), h = (r.y + r.z).sqr() - (b + c), i = e - b, j = r.x.sqr(), e2 = e.sqr();
    r.x = a * (b - f);
    r.y = g.sqr() - (e2.dbl() + e2);
    r.z = b * h;
    return Ell{h.neg(), j.dbl() + j, i};
}
static Ell add_step(Hom& r, const G2A& q) {
    Fq2 theta = r.y - q.y * r.z, lambda = r.x - q.x * r.z;
    Fq2 c = theta.sqr(), d = lambda.sqr(), e = lambda * d, f = r.z * c, g = r.x * d;
    Fq2 h = e + f - g.dbl();
    r.x = lambda * h;
    r.y = theta * (g - h) - e * r.y;
    r.z = r.z * e;
    return Ell{lambda, theta.neg(), theta * q.x - lambda * q.y};
}
static G2A mul_by_char(const G2A& q) { return G2A{conj2(q.x) * K().twist_x, conj2(q.y) * K().twist_y}; }

static std::vector<Ell> g2_prepare(const G2A& q) {
    std::vector<Ell> out;
    Hom r{q.x, q.y, Fq2::one()};
    G2A nq{q.x, q.y.neg()};
    for (int i = 63; i >= 0; i--) {
        out.push_back(double_step(r));
        if (ATE[i] == 1) out.push_back(add_step(r, q));
        else if (ATE[i] == -1) out.push_back(add_step(r, nq));
    }
    G2A q1 = mul_by_char(q), q2 = mul_by_char(q1);
    q2.y = q2.y.neg();
    out.push_back(add_step(r, q1));
    out.push_back(add_step(r, q2));
    return out;
}

// f * (a0 + (d0 + d1 v) w)
static Fq12 mul_by_034(const Fq12& f, const Fq2& a0, const Fq2& d0, const Fq2& d1) {
    Fq6 s1{d0, d1, Fq2::zero()};
    Fq6 a{f.c0.c0 * a0, f.c0.c1 * a0, f.c0.c2 * a0};
    Fq6 b = f.c1 * s1;
    Fq6 e = (f.c0 + f.c1) * Fq6{a0 + d0, d1, Fq2::zero()};
    return Fq12{a + b.mul_v(), e - a - b};
}
static void ell(Fq12& f, const Ell& c, const G1A& p) { f = mul_by_034(f, scale2(c.c0, p.y), scale2(c.c1, p.x), c.c2); }

struct Pair {
    G1A p;
    const std::vector<Ell>* coeffs;
};
static Fq12 multi_miller_loop(const std::vector<Pair>& pairs) {
    Fq12 f = Fq12::one();
    size_t idx = 0;
    for (int i = 63; i >= 0; i--) {
        if (i != 63) f = f.sqr();
        for (const Pair& pr : pairs) ell(f, (*pr.coeffs)[idx], pr.p);
        idx++;
        if (ATE[i] != 0) {
            for (const Pair& pr : pairs) ell(f, (*pr.coeffs)[idx], pr.p);
            idx++;
        }
    }
    for (int k = 0; k < 2; k++, idx++)
        for (const Pair& pr : pairs) ell(f, (*pr.coeffs)[idx], pr.p);
    return f;
}

static Fq12 exp_by_neg_x(const Fq12& f) {
    Fq12 r = f;
    for (int i = 61; i >= 0; i--) {
        r = cyclotomic_sqr(r);
        if ((BN_X >> i) & 1) r = r * f;
    }
    return r.conj();
}
static bool final_exponentiation(const Fq12& f, Fq12& out) {
    if (f.is_zero()) return false;
    Fq12 r = f.conj() * f.inverse();
    r = frobenius(r, 2) * r;
    Fq12 y0 = exp_by_neg_x(r), y1 = cyclotomic_sqr(y0), y2 = cyclotomic_sqr(y1), y3 = y2 * y1;
    Fq12 y4 = exp_by_neg_x(y3), y5 = cyclotomic_sqr(y4), y6 = exp_by_neg_x(y5).conj();
    y3 = y3.conj();
    Fq12 y7 = y6 * y4, y8 = y7 * y3, y9 = y8 * y1, y10 = y8 * y4, y11 = y10 * r;
    Fq12 y13 = frobenius(y9, 1) * y11, y14 = frobenius(y8, 2) * y13, y15 = frobenius(r.conj() * y9, 3);
    out = y15 * y14;
    return true;
}

struct Pvk {
    Fq12 alpha_beta;
    std::vector<Ell> neg_gamma, neg_delta;
    bool gamma_inf, delta_inf;
};
struct oc_vk {
    const G1A* alpha_g1;
    const G2A* beta_g2;
    const G2A* gamma_g2;
    const G2A* delta_g2;
    const G1A* gamma_abc_g1;
    size_t gamma_abc_len;
};
static Fq12 pairing(const G1A& p, const G2A& q) {
    Fq12 out = Fq12::one();
    if (p.is_inf() || q.is_inf()) return out;
    std::vector<Ell> c = g2_prepare(q);
    final_exponentiation(multi_miller_loop({Pair{p, &c}}), out);
    return out;
}
static Pvk prepare_vk(const oc_vk* vk) {  // verifier.rs:13-20
    Pvk k;
    k.alpha_beta = pairing(*vk->alpha_g1, *vk->beta_g2);
    k.gamma_inf = vk->gamma_g2->is_inf();
    k.delta_inf = vk->delta_g2->is_inf();
    if (!k.gamma_inf) k.neg_gamma = g2_prepare(G2A{vk->gamma_g2->x, vk->gamma_g2->y.neg()});
    if (!k.delta_inf) k.neg_delta = g2_prepare(G2A{vk->delta_g2->x, vk->delta_g2->y.neg()});
    return k;
}
static G1A prepare_inputs(const oc_vk* vk, const Fr* x) {  // verifier.rs:25-39
    G1J acc = G1J::from_affine(vk->gamma_abc_g1[0]);
    for (size_t i = 0; i + 1 < vk->gamma_abc_len; i++) {
        Fr c = x[i].from_mont();
        acc.add(G1J::from_affine(vk->gamma_abc_g1[i + 1]).mul(c.l));
    }
    return acc.to_affine();
}
// verifier.rs:44-65; returns 1 / 0 / 2 (UnexpectedIdentity)
static int verify_prepared(const Pvk& k, const G1A& a, const G2A& b, const G1A& c, const G1A& prepared) {
    std::vector<Ell> bc;
    std::vector<Pair> pairs;
    if (!a.is_inf() && !b.is_inf()) {
        bc = g2_prepare(b);
        pairs.push_back(Pair{a, &bc});
    }
    if (!prepared.is_inf() && !k.gamma_inf) pairs.push_back(Pair{prepared, &k.neg_gamma});
    if (!c.is_inf() && !k.delta_inf) pairs.push_back(Pair{c, &k.neg_delta});
    Fq12 t;
    if (!final_exponentiation(multi_miller_loop(pairs), t)) return 2;
    return t == k.alpha_beta ? 1 : 0;
}

}  // namespace pairing_oracle

extern "C" {

// gt_out: n x 12 Montgomery Fq in ark-serialize order
void oc_pairing(const uint64_t* g1, const uint64_t* g2, size_t n, uint64_t* gt_out, int threads) {
    using namespace pairing_oracle;
    static_assert(sizeof(Fq12) == 384, "Fq12 layout");
    parallel_for(n, threads > 0 ? threads : (int)std::thread::hardware_concurrency(), [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            Fq12 e = pairing(((const G1A*)g1)[i], ((const G2A*)g2)[i]);
            memcpy(gt_out + 48 * i, &e, sizeof(e));
        }
    });
}
void oc_prepare_vk(const pairing_oracle::oc_vk* vk, uint64_t* alpha_beta_out) {
    pairing_oracle::Pvk k = pairing_oracle::prepare_vk(vk);
    memcpy(alpha_beta_out, &k.alpha_beta, 384);
}
void oc_prepare_inputs(const pairing_oracle::oc_vk* vk, const uint64_t* inputs, size_t n, uint64_t* out, int threads) {
    size_t k = vk->gamma_abc_len - 1;
    parallel_for(n, threads > 0 ? threads : (int)std::thread::hardware_concurrency(), [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) ((G1A*)out)[i] = pairing_oracle::prepare_inputs(vk, (const Fr*)inputs + i * k);
    });
}
// proofs: n x {a (8 u64), b (16), c (8)} Montgomery affine, infinity = zeros; verdict[i] = 1 / 0 / 2.  Returns the seconds spent
// in the per-proof part (prepare_inputs + verify), the key preparation excluded.
double oc_verify(const pairing_oracle::oc_vk* vk, const oc_proof* proofs, const uint64_t* inputs, size_t n, uint8_t* verdict, int threads) {
    using namespace pairing_oracle;
    Pvk pvk = prepare_vk(vk);
    size_t k = vk->gamma_abc_len - 1;
    double t0 = now_s();
    parallel_for(n, threads > 0 ? threads : (int)std::thread::hardware_concurrency(), [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            G1A pi = prepare_inputs(vk, (const Fr*)inputs + i * k);
            verdict[i] = (uint8_t)verify_prepared(pvk, proofs[i].a, proofs[i].b, proofs[i].c, pi);
        }
    });
    return now_s() - t0;
}

}  // extern "C"
