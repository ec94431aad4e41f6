// ark_groth16_b200.hpp -- C++17 host-side mirror of the ark-groth16 API surface Crescent uses, sitting directly on the
// C ABI of libg16b200.so (include/g16_b200.h).  The reference host code is Rust (compiled); this image has no Rust
// toolchain, so this header is the compiled host side: same names, argument meaning and error behaviour as the fork.
//
//   Groth16<QAP>::create_proof_with_reduction_and_matrices(pk, r, s, matrices, num_inputs, num_constraints, full_assignment)
//                                                          -> forks/groth16/src/prover.rs:26-51
//   Groth16<QAP>::create_proof_with_reduction(circuit, pk, r, s)          -> prover.rs:177-221 (CircomCircuit, circuit.rs:28-87)
//   Groth16<QAP>::create_random_proof_with_reduction / _no_zk              -> prover.rs:142-172
//   Groth16<QAP>::witness_map_from_matrices                                -> r1cs_to_qap.rs:150-213 / circom qap.rs:25-90
//   Proof / VerifyingKey / ProvingKey (+ ark-serialize 0.4 layout)         -> data_structures.rs:7-14,31-44,101-118
//   R1CSFile / R1CS / CircomCircuit                                        -> forks/circom-compat/src/circom/r1cs_reader.rs:54-256
//   StdRng / test_rng / Fr::rand                                           -> rand 0.8 / ark-std 0.4 / ark-ff 0.4 (see rng.py)
//   ShardedGroth16 (MSMs sharded over the GPUs of one box, one host thread per GPU)  -> SURVEY 8e
//
// All field and curve arithmetic of the proof runs in the CUDA library.  The host only marshals: limb packing, the
// Montgomery <-> canonical conversion of a handful of elements (r, s, the six proof coordinates: what arkworks'
// `into_bigint` does inside `serialize`), flag bits, file parsing.  There is no CPU prover and no fallback: every
// compute call fails with G16_ERR_NO_DEVICE when no GPU is visible.
#ifndef ARK_GROTH16_B200_HPP
#define ARK_GROTH16_B200_HPP

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "g16_b200.h"

namespace ark_groth16_b200 {

// ---- errors ---------------------------------------------------------------------------------------------------------------
// R1CSResult<T> = Result<T, SynthesisError>: the one SynthesisError raised on the path is thrown as SynthesisError; every
// other failure (bad sizes, CUDA, out of memory) is a Panic, matching the `.unwrap()`s at creds/src/lib.rs:283.
struct SynthesisError : std::runtime_error {
    enum Kind { PolynomialDegreeTooLarge, AssignmentMissing, MalformedVerifyingKey, UnexpectedIdentity } kind;
    SynthesisError(Kind k, const std::string& m) : std::runtime_error(m), kind(k) {}
};
struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct SerializationError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- 256-bit marshalling helpers (host) ---------------------------------------------------------------------------------------
using Limbs = std::array<uint64_t, 4>;

namespace detail {
constexpr Limbs kFrModulus = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
constexpr Limbs kFqModulus = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
constexpr uint64_t kFrInv = 0xc2e1f593efffffffull;  // -r^-1 mod 2^64
constexpr uint64_t kFqInv = 0x87d20782e4866389ull;  // -q^-1 mod 2^64
constexpr Limbs kFrR2 = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
constexpr Limbs kFqR2 = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};

inline int cmp(const Limbs& a, const Limbs& b) {
    for (int i = 3; i >= 0; i--)
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}
inline bool is_zero(const Limbs& a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
inline Limbs sub(const Limbs& a, const Limbs& b) {  // a - b, a >= b
    Limbs r;
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 d = (unsigned __int128)a[i] - b[i] - (uint64_t)borrow;
        r[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    return r;
}
inline Limbs add_mod(const Limbs& a, const Limbs& b, const Limbs& p) {  // a, b < p < 2^254: no overflow of 256 bits
    Limbs r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return cmp(r, p) >= 0 ? sub(r, p) : r;
}
// Montgomery product a*b/2^256 mod p (CIOS on 64-bit limbs).  Used for single elements only (r, s, proof coordinates).
inline Limbs mont_mul(const Limbs& a, const Limbs& b, const Limbs& p, uint64_t inv) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (unsigned __int128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * inv;
        c = ((unsigned __int128)m * p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (unsigned __int128)m * p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Limbs r = {t[0], t[1], t[2], t[3]};
    if (t[4] || cmp(r, p) >= 0) r = sub(r, p);
    return r;
}
inline void put_le(std::vector<uint8_t>& out, const Limbs& v) {
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 8; k++) out.push_back((uint8_t)(v[i] >> (8 * k)));
}
inline Limbs get_le(const uint8_t* p) {
    Limbs v;
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0;
        for (int k = 7; k >= 0; k--) w = (w << 8) | p[8 * i + k];
        v[i] = w;
    }
    return v;
}
inline Limbs from_hex(std::string s) {
    if (s.rfind("0x", 0) == 0 || s.rfind("0X", 0) == 0) s = s.substr(2);
    if (s.size() > 64) throw std::invalid_argument("hex value wider than 256 bits");
    s = std::string(64 - s.size(), '0') + s;
    Limbs v;
    for (int i = 0; i < 4; i++) v[3 - i] = std::stoull(s.substr(16 * i, 16), nullptr, 16);
    return v;
}
inline std::string to_hex(const Limbs& v) {
    static const char* d = "0123456789abcdef";
    std::string s;
    for (int i = 3; i >= 0; i--)
        for (int k = 15; k >= 0; k--) s.push_back(d[(v[i] >> (4 * k)) & 15]);
    return s;
}
}  // namespace detail

// Prime-field element held as arkworks holds it: 4 x u64 little-endian limbs in Montgomery form.
template <int FIELD>
struct Fp {
    Limbs v{};  // Montgomery representation
    static const Limbs& modulus() { return FIELD == G16_FIELD_FR ? detail::kFrModulus : detail::kFqModulus; }
    static uint64_t inv() { return FIELD == G16_FIELD_FR ? detail::kFrInv : detail::kFqInv; }
    static const Limbs& r2() { return FIELD == G16_FIELD_FR ? detail::kFrR2 : detail::kFqR2; }
    static Fp new_unchecked(const Limbs& mont) { return Fp{mont}; }
    // PrimeField::from_bigint: canonical integer (< modulus) -> element; nullopt when out of range
    static std::optional<Fp> from_bigint(const Limbs& canonical) {
        if (detail::cmp(canonical, modulus()) >= 0) return std::nullopt;
        return Fp{detail::mont_mul(canonical, r2(), modulus(), inv())};
    }
    static Fp from_u64(uint64_t x) { return *from_bigint(Limbs{x, 0, 0, 0}); }
    static Fp zero() { return Fp{}; }
    static Fp one() { return from_u64(1); }
    Limbs into_bigint() const { return detail::mont_mul(v, Limbs{1, 0, 0, 0}, modulus(), inv()); }
    bool is_zero() const { return detail::is_zero(v); }
    bool operator==(const Fp& o) const { return v == o.v; }
    bool operator!=(const Fp& o) const { return !(v == o.v); }
    // UniformRand for Fp (ark-ff 0.4): four next_u64 draws (limb 0 first), the top limb masked to the modulus' bit length
    // (254), rejection above the modulus; the accepted integer IS the Montgomery representation.
    template <class Rng>
    static Fp rand(Rng& rng) {
        for (;;) {
            Limbs x;
            for (int i = 0; i < 4; i++) x[i] = rng.next_u64();
            x[3] &= (~0ull) >> 2;
            if (detail::cmp(x, modulus()) < 0) return Fp{x};
        }
    }
};
using Fr = Fp<G16_FIELD_FR>;
using Fq = Fp<G16_FIELD_FQ>;
struct Fq2 {
    Fq c0, c1;
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
};

// short_weierstrass::Affine {x, y, infinity}
struct G1Affine {
    Fq x, y;
    bool infinity = true;
    static G1Affine identity() { return G1Affine{}; }
    bool is_zero() const { return infinity; }
    bool operator==(const G1Affine& o) const { return infinity == o.infinity && (infinity || (x == o.x && y == o.y)); }
};
struct G2Affine {
    Fq2 x, y;
    bool infinity = true;
    static G2Affine identity() { return G2Affine{}; }
    bool is_zero() const { return infinity; }
    bool operator==(const G2Affine& o) const { return infinity == o.infinity && (infinity || (x == o.x && y == o.y)); }
};

// ---- ark-serialize 0.4 (little-endian; SW flags in the two top bits of the last byte) ------------------------------------------------
enum class Compress { Yes, No };
namespace detail {
constexpr uint8_t kFlagNegative = 0x80, kFlagInfinity = 0x40;
// "y is the larger of {y, -y}" on canonical integers; Fq2 orders by c1 then c0
inline bool fq_is_negative(const Limbs& y) { return !is_zero(y) && cmp(y, sub(kFqModulus, y)) > 0; }
inline bool fq2_is_negative(const Limbs& y0, const Limbs& y1) {
    Limbs n0 = is_zero(y0) ? y0 : sub(kFqModulus, y0), n1 = is_zero(y1) ? y1 : sub(kFqModulus, y1);
    int c = cmp(y1, n1);
    return c ? c > 0 : cmp(y0, n0) > 0;
}
}  // namespace detail

inline void serialize_g1(const G1Affine& p, Compress c, std::vector<uint8_t>& out) {
    size_t len = c == Compress::Yes ? 32 : 64;
    if (p.infinity) {
        out.insert(out.end(), len, 0);
        out.back() |= detail::kFlagInfinity;
        return;
    }
    Limbs y = p.y.into_bigint();
    detail::put_le(out, p.x.into_bigint());
    if (c == Compress::No) detail::put_le(out, y);
    if (detail::fq_is_negative(y)) out.back() |= detail::kFlagNegative;
}
inline void serialize_g2(const G2Affine& p, Compress c, std::vector<uint8_t>& out) {
    size_t len = c == Compress::Yes ? 64 : 128;
    if (p.infinity) {
        out.insert(out.end(), len, 0);
        out.back() |= detail::kFlagInfinity;
        return;
    }
    Limbs y0 = p.y.c0.into_bigint(), y1 = p.y.c1.into_bigint();
    detail::put_le(out, p.x.c0.into_bigint());
    detail::put_le(out, p.x.c1.into_bigint());
    if (c == Compress::No) {
        detail::put_le(out, y0);
        detail::put_le(out, y1);
    }
    if (detail::fq2_is_negative(y0, y1)) out.back() |= detail::kFlagNegative;
}

// Proof{a, b, c}: data_structures.rs:7-14.  serialize_compressed = 128 bytes, serialize_uncompressed = 256 bytes (what
// creds/src/utils.rs:140-152 persists).
struct Proof {
    G1Affine a;
    G2Affine b;
    G1Affine c;
    std::vector<uint8_t> serialize(Compress m) const {
        std::vector<uint8_t> out;
        serialize_g1(a, m, out);
        serialize_g2(b, m, out);
        serialize_g1(c, m, out);
        return out;
    }
    std::vector<uint8_t> serialize_compressed() const { return serialize(Compress::Yes); }
    std::vector<uint8_t> serialize_uncompressed() const { return serialize(Compress::No); }
    bool operator==(const Proof& o) const { return a == o.a && b == o.b && c == o.c; }
    // CanonicalDeserialize::deserialize_uncompressed_unchecked of the 256 bytes Crescent persists (creds/src/utils.rs:140-152):
    // canonical little-endian coordinates, flags in the top bits of each point's last byte; converted to Montgomery on the host
    static Proof deserialize_uncompressed_unchecked(const uint8_t* buf, size_t len) {
        if (len != 256) throw SerializationError("Proof: expected 256 bytes");
        auto fq = [&](size_t off, bool last) {
            Limbs v = detail::get_le(buf + off);
            if (last) v[3] &= (~0ull) >> 2;
            auto f = Fq::from_bigint(v);
            if (!f) throw SerializationError("Proof: coordinate not below the modulus");
            return *f;
        };
        auto inf = [&](size_t last_byte) { return (buf[last_byte] & detail::kFlagInfinity) != 0; };
        Proof r;
        if (!inf(63)) r.a = G1Affine{fq(0, false), fq(32, true), false};
        if (!inf(191)) r.b = G2Affine{Fq2{fq(64, false), fq(96, false)}, Fq2{fq(128, false), fq(160, true)}, false};
        if (!inf(255)) r.c = G1Affine{fq(192, false), fq(224, true), false};
        return r;
    }
    g16_proof to_abi() const {
        g16_proof p{};
        auto put = [](uint64_t* w, const Fq& f) { std::memcpy(w, f.v.data(), 32); };
        if (!a.infinity) put(p.a, a.x), put(p.a + 4, a.y);
        if (!b.infinity) put(p.b, b.x.c0), put(p.b + 4, b.x.c1), put(p.b + 8, b.y.c0), put(p.b + 12, b.y.c1);
        if (!c.infinity) put(p.c, c.x), put(p.c + 4, c.y);
        p.a_inf = a.infinity, p.b_inf = b.infinity, p.c_inf = c.infinity;
        return p;
    }
    static Proof from_abi(const g16_proof& p) {
        auto fq = [](const uint64_t* w) { return Fq{Limbs{w[0], w[1], w[2], w[3]}}; };
        Proof r;
        if (!p.a_inf) r.a = G1Affine{fq(p.a), fq(p.a + 4), false};
        if (!p.b_inf) r.b = G2Affine{Fq2{fq(p.b), fq(p.b + 4)}, Fq2{fq(p.b + 8), fq(p.b + 12)}, false};
        if (!p.c_inf) r.c = G1Affine{fq(p.c), fq(p.c + 4), false};
        return r;
    }
};

// Packed point vectors in the layout of the C ABI: G1 = x || y (8 x u64), G2 = x.c0 || x.c1 || y.c0 || y.c1 (16 x u64),
// infinity = all zero.  `encoding` says whether the words are Montgomery limbs (arkworks in memory) or canonical integers
// (arkworks serialised); the library converts canonical input on the GPU.
struct PointVec {
    std::vector<uint64_t> w;
    int words = 8;
    size_t size() const { return w.size() / words; }
    const uint64_t* data() const { return w.data(); }
};

// VerifyingKey (data_structures.rs:31-44; delta_g1 is the fork's extra field) and ProvingKey (:101-118), in serialisation order.
struct VerifyingKey {
    PointVec alpha_g1{{}, 8}, beta_g2{{}, 16}, gamma_g2{{}, 16}, delta_g1{{}, 8}, delta_g2{{}, 16}, gamma_abc_g1{{}, 8};
};
struct ProvingKey {
    VerifyingKey vk;
    PointVec beta_g1{{}, 8}, delta_g1{{}, 8}, a_query{{}, 8}, b_g1_query{{}, 8}, b_g2_query{{}, 16}, h_query{{}, 8}, l_query{{}, 8};
    int encoding = G16_ENC_MONTGOMERY;
    // infinity + sign flags of every point as read (kept so that serialize_uncompressed reproduces the input bytes)
    std::vector<uint8_t> flags;

    // CanonicalDeserialize::deserialize_uncompressed_unchecked of arkworks' ProvingKey bytes -- Crescent's
    // cache/prover_params.bin (creds/src/utils.rs:179-189).  Coordinates stay canonical words (flag bits cleared, infinity
    // -> zeros): no per-point arithmetic on the host.
    static ProvingKey deserialize_uncompressed_unchecked(const uint8_t* buf, size_t len) {
        ProvingKey pk;
        pk.encoding = G16_ENC_CANONICAL;
        size_t off = 0;
        auto need = [&](size_t n) {
            if (len - off < n) throw SerializationError("ProvingKey: unexpected end of input");
        };
        auto points = [&](PointVec& dst, size_t count) {
            size_t bytes = (size_t)dst.words * 8;
            if (count > (len - off) / bytes) throw SerializationError("ProvingKey: vector length exceeds the input");
            dst.w.resize(count * dst.words);
            std::memcpy(dst.w.data(), buf + off, count * bytes);  // little-endian host
            off += count * bytes;
            for (size_t i = 0; i < count; i++) {
                uint64_t& top = dst.w[i * dst.words + dst.words - 1];
                uint8_t f = (uint8_t)(top >> 62);
                pk.flags.push_back(f);
                top &= (~0ull) >> 2;
                if (f & 1) std::fill_n(dst.w.begin() + i * dst.words, dst.words, 0ull);  // infinity flag (bit 6 of the last byte)
            }
        };
        auto vec_len = [&]() {
            need(8);
            uint64_t n;
            std::memcpy(&n, buf + off, 8);
            off += 8;
            return (size_t)n;
        };
        points(pk.vk.alpha_g1, 1);
        points(pk.vk.beta_g2, 1);
        points(pk.vk.gamma_g2, 1);
        points(pk.vk.delta_g1, 1);
        points(pk.vk.delta_g2, 1);
        points(pk.vk.gamma_abc_g1, vec_len());
        points(pk.beta_g1, 1);
        points(pk.delta_g1, 1);
        points(pk.a_query, vec_len());
        points(pk.b_g1_query, vec_len());
        points(pk.b_g2_query, vec_len());
        points(pk.h_query, vec_len());
        points(pk.l_query, vec_len());
        if (off != len) throw SerializationError("ProvingKey: trailing bytes");
        return pk;
    }
    static ProvingKey deserialize_uncompressed_unchecked(const std::vector<uint8_t>& b) {
        return deserialize_uncompressed_unchecked(b.data(), b.size());
    }
    // CanonicalSerialize::serialize_uncompressed (only for keys read from canonical bytes: the flags are replayed)
    std::vector<uint8_t> serialize_uncompressed() const {
        if (encoding != G16_ENC_CANONICAL) throw SerializationError("ProvingKey: only canonical-encoded keys re-serialise on the host");
        std::vector<uint8_t> out;
        size_t fi = 0;
        auto points = [&](const PointVec& src) {
            size_t at = out.size(), bytes = src.w.size() * 8;
            out.resize(at + bytes);
            std::memcpy(out.data() + at, src.w.data(), bytes);
            for (size_t i = 0; i < src.size(); i++) out[at + (i + 1) * src.words * 8 - 1] |= (uint8_t)(flags.at(fi++) << 6);
        };
        auto vec = [&](const PointVec& src) {
            uint64_t n = src.size();
            for (int k = 0; k < 8; k++) out.push_back((uint8_t)(n >> (8 * k)));
            points(src);
        };
        points(vk.alpha_g1);
        points(vk.beta_g2);
        points(vk.gamma_g2);
        points(vk.delta_g1);
        points(vk.delta_g2);
        vec(vk.gamma_abc_g1);
        points(beta_g1);
        points(delta_g1);
        vec(a_query);
        vec(b_g1_query);
        vec(b_g2_query);
        vec(h_query);
        vec(l_query);
        return out;
    }
    g16_pk_view view() const {
        g16_pk_view v{};
        v.a_query = a_query.data(), v.a_len = a_query.size();
        v.b_g1_query = b_g1_query.data(), v.b_g1_len = b_g1_query.size();
        v.b_g2_query = b_g2_query.data(), v.b_g2_len = b_g2_query.size();
        v.h_query = h_query.data(), v.h_len = h_query.size();
        v.l_query = l_query.data(), v.l_len = l_query.size();
        v.alpha_g1 = vk.alpha_g1.data(), v.beta_g1 = beta_g1.data(), v.delta_g1 = delta_g1.data();
        v.beta_g2 = vk.beta_g2.data(), v.delta_g2 = vk.delta_g2.data();
        v.encoding = encoding;
        return v;
    }
};

// ---- ConstraintMatrices<Fr> (ark-relations) flattened to CSR ------------------------------------------------------------
struct Csr {
    std::vector<uint64_t> row_ptr{0};
    std::vector<uint32_t> col;
    std::vector<uint64_t> val;  // 4 words per entry
    size_t num_non_zero() const { return col.size(); }
};
struct ConstraintMatrices {
    size_t num_instance_variables = 0, num_witness_variables = 0, num_constraints = 0;
    Csr a, b, c;
    int encoding = G16_ENC_MONTGOMERY;  // of the coefficient words
    size_t a_num_non_zero() const { return a.num_non_zero(); }
    size_t b_num_non_zero() const { return b.num_non_zero(); }
    size_t c_num_non_zero() const { return c.num_non_zero(); }
    using Row = std::vector<std::pair<Fr, size_t>>;  // (coeff, column): one Vec per constraint, as in ark-relations
    static ConstraintMatrices from_rows(size_t num_instance, size_t num_witness, const std::vector<Row>& ra, const std::vector<Row>& rb,
                                        const std::vector<Row>& rc) {
        ConstraintMatrices m;
        m.num_instance_variables = num_instance, m.num_witness_variables = num_witness, m.num_constraints = ra.size();
        if (rb.size() != ra.size() || rc.size() != ra.size()) throw Panic(G16_ERR_BAD_ARG, "a, b, c must have one row per constraint");
        auto flat = [](const std::vector<Row>& rows, Csr& o) {
            for (auto& row : rows) {
                for (auto& e : row) {
                    o.col.push_back((uint32_t)e.second);
                    o.val.insert(o.val.end(), e.first.v.begin(), e.first.v.end());
                }
                o.row_ptr.push_back(o.col.size());
            }
        };
        flat(ra, m.a), flat(rb, m.b), flat(rc, m.c);
        return m;
    }
    g16_r1cs_view view() const {
        g16_r1cs_view v{};
        v.num_constraints = num_constraints, v.num_instance = num_instance_variables;
        v.num_wires = num_instance_variables + num_witness_variables;
        const Csr* m[3] = {&a, &b, &c};
        for (int k = 0; k < 3; k++) v.row_ptr[k] = m[k]->row_ptr.data(), v.col[k] = m[k]->col.data(), v.val[k] = m[k]->val.data();
        v.encoding = encoding;
        return v;
    }
};

// ---- iden3 .r1cs reader (forks/circom-compat/src/circom/r1cs_reader.rs:54-256; worked example at :266-344) --------------------------
struct R1CSFile {
    uint32_t version = 0, field_size = 0, n_wires = 0, n_pub_out = 0, n_pub_in = 0, n_prv_in = 0, n_constraints = 0;
    uint64_t n_labels = 0;
    std::vector<uint64_t> wire_mapping;
    using LC = std::vector<std::pair<uint32_t, Limbs>>;  // (wire, canonical coefficient)
    std::vector<std::array<LC, 3>> constraints;

    static R1CSFile read(const uint8_t* d, size_t len) {
        static const uint8_t kPrime[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                                           0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
        size_t off = 0;
        auto need = [&](size_t at, size_t n) {
            if (at > len || len - at < n) throw SerializationError("r1cs: unexpected end of file");
        };
        auto u32 = [&](size_t at) {
            need(at, 4);
            uint32_t v;
            std::memcpy(&v, d + at, 4);
            return v;
        };
        auto u64 = [&](size_t at) {
            need(at, 8);
            uint64_t v;
            std::memcpy(&v, d + at, 8);
            return v;
        };
        need(0, 12);
        if (std::memcmp(d, "r1cs", 4) != 0) throw SerializationError("Invalid magic number");
        R1CSFile f;
        f.version = u32(4);
        if (f.version != 1) throw SerializationError("Unsupported version");
        uint32_t nsec = u32(8);
        off = 12;
        std::map<uint32_t, std::pair<size_t, size_t>> secs;
        for (uint32_t i = 0; i < nsec; i++) {
            uint32_t ty = u32(off);
            uint64_t sz = u64(off + 4);
            off += 12;
            need(off, sz);
            secs[ty] = {off, (size_t)sz};
            off += sz;
        }
        const char* names[] = {"", "header", "constraint", "wire2label"};
        for (uint32_t ty = 1; ty <= 3; ty++)
            if (!secs.count(ty)) throw SerializationError(std::string("No section offset for ") + names[ty] + " type found");
        size_t ho = secs[1].first;
        f.field_size = u32(ho);
        if (f.field_size != 32) throw SerializationError("This parser only supports 32-byte fields");
        if (secs[1].second != 32 + f.field_size) throw SerializationError("Invalid header section size");
        need(ho + 4, 32);
        if (std::memcmp(d + ho + 4, kPrime, 32) != 0) throw SerializationError("This parser only supports bn256");
        f.n_wires = u32(ho + 36), f.n_pub_out = u32(ho + 40), f.n_pub_in = u32(ho + 44), f.n_prv_in = u32(ho + 48);
        f.n_labels = u64(ho + 52), f.n_constraints = u32(ho + 60);
        size_t p = secs[2].first;
        f.constraints.resize(f.n_constraints);
        for (uint32_t i = 0; i < f.n_constraints; i++)
            for (int k = 0; k < 3; k++) {
                uint32_t nv = u32(p);
                p += 4;
                need(p, (size_t)nv * 36);
                LC& lc = f.constraints[i][k];
                lc.reserve(nv);
                for (uint32_t j = 0; j < nv; j++, p += 36) lc.emplace_back(u32(p), detail::get_le(d + p + 4));
            }
        if (secs[3].second != (size_t)f.n_wires * 8) throw SerializationError("Invalid map section size");
        f.wire_mapping.resize(f.n_wires);
        for (uint32_t i = 0; i < f.n_wires; i++) f.wire_mapping[i] = u64(secs[3].first + 8 * (size_t)i);
        if (f.n_wires && f.wire_mapping[0] != 0) throw SerializationError("Wire 0 should always be mapped to 0");
        return f;
    }
    static R1CSFile read(const std::vector<uint8_t>& b) { return read(b.data(), b.size()); }
};

// R1CS (circom-compat r1cs_reader.rs:17-44): num_inputs = 1 + n_pub_in + n_pub_out, num_aux = n_wires - num_inputs,
// wire_mapping forced to None by the builder (builder.rs:64), so a column is the circom wire index.
struct R1CS {
    size_t num_inputs = 0, num_aux = 0, num_variables = 0;
    std::vector<std::array<R1CSFile::LC, 3>> constraints;
    static R1CS from_file(R1CSFile&& f) {
        R1CS r;
        r.num_inputs = 1 + (size_t)f.n_pub_in + f.n_pub_out;
        r.num_variables = f.n_wires;
        r.num_aux = r.num_variables - r.num_inputs;
        r.constraints = std::move(f.constraints);
        return r;
    }
    // What ConstraintSystem::to_matrices() yields after CircomCircuit::generate_constraints (circuit.rs:28-87) and LC
    // inlining: per row, duplicate wires summed, zero coefficients dropped, columns ascending.  Built ONCE per circuit
    // (SURVEY 8f-1); the coefficients stay canonical words (the library converts them on the GPU).
    ConstraintMatrices to_matrices() const {
        ConstraintMatrices m;
        m.num_instance_variables = num_inputs, m.num_witness_variables = num_aux, m.num_constraints = constraints.size();
        m.encoding = G16_ENC_CANONICAL;
        Csr* out[3] = {&m.a, &m.b, &m.c};
        std::vector<std::pair<uint32_t, Limbs>> acc;
        for (auto& con : constraints)
            for (int k = 0; k < 3; k++) {
                acc.assign(con[k].begin(), con[k].end());
                std::stable_sort(acc.begin(), acc.end(), [](auto& x, auto& y) { return x.first < y.first; });
                for (size_t i = 0; i < acc.size();) {
                    if (acc[i].first >= num_variables) throw SerializationError("wire index out of range");
                    if (detail::cmp(acc[i].second, detail::kFrModulus) >= 0) throw SerializationError("coefficient not reduced");
                    Limbs sum = acc[i].second;
                    size_t j = i + 1;
                    for (; j < acc.size() && acc[j].first == acc[i].first; j++) {
                        if (detail::cmp(acc[j].second, detail::kFrModulus) >= 0) throw SerializationError("coefficient not reduced");
                        sum = detail::add_mod(sum, acc[j].second, detail::kFrModulus);
                    }
                    if (!detail::is_zero(sum)) {
                        out[k]->col.push_back(acc[i].first);
                        out[k]->val.insert(out[k]->val.end(), sum.begin(), sum.end());
                    }
                    i = j;
                }
                out[k]->row_ptr.push_back(out[k]->col.size());
            }
        return m;
    }
};

// CircomCircuit {r1cs, witness} (circuit.rs:12-16): witness = the full wire assignment, wire 0 = 1.
struct CircomCircuit {
    std::shared_ptr<const R1CS> r1cs;
    std::optional<std::vector<Fr>> witness;
};

// ---- rand 0.8 StdRng (ChaCha12 behind rand_core's BlockRng) -- see crescent_credentials_b200/rng.py for the provenance -----------------
class ChaChaRng {
  public:
    explicit ChaChaRng(const std::array<uint8_t, 32>& seed, int rounds = 12, uint64_t stream = 0) : rounds_(rounds), stream_(stream) {
        for (int i = 0; i < 8; i++) key_[i] = (uint32_t)seed[4 * i] | (uint32_t)seed[4 * i + 1] << 8 | (uint32_t)seed[4 * i + 2] << 16 | (uint32_t)seed[4 * i + 3] << 24;
    }
    static ChaChaRng from_seed(const std::array<uint8_t, 32>& seed) { return ChaChaRng(seed); }
    // rand_core 0.6 SeedableRng::seed_from_u64: eight PCG32 outputs, little-endian
    static ChaChaRng seed_from_u64(uint64_t state) {
        std::array<uint8_t, 32> seed{};
        for (int i = 0; i < 8; i++) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            for (int k = 0; k < 4; k++) seed[4 * i + k] = (uint8_t)(x >> (8 * k));
        }
        return ChaChaRng(seed);
    }
    uint32_t next_u32() {
        if (index_ >= kBuf) {
            generate();
            index_ = 0;
        }
        return buf_[index_++];
    }
    uint64_t next_u64() {
        size_t i = index_;
        if (i + 1 < kBuf) {
            index_ += 2;
            return buf_[i] | (uint64_t)buf_[i + 1] << 32;
        }
        if (i >= kBuf) {
            generate();
            index_ = 2;
            return buf_[0] | (uint64_t)buf_[1] << 32;
        }
        uint64_t lo = buf_[kBuf - 1];  // one word left: low half; the first word of the next buffer is the high half
        generate();
        index_ = 1;
        return lo | (uint64_t)buf_[0] << 32;
    }
    static void block(const uint32_t in[16], int rounds, uint32_t out[16]) {
        uint32_t w[16];
        std::memcpy(w, in, sizeof w);
        auto rotl = [](uint32_t x, int n) { return (x << n) | (x >> (32 - n)); };
        auto qr = [&](int a, int b, int c, int d) {
            w[a] += w[b], w[d] = rotl(w[d] ^ w[a], 16);
            w[c] += w[d], w[b] = rotl(w[b] ^ w[c], 12);
            w[a] += w[b], w[d] = rotl(w[d] ^ w[a], 8);
            w[c] += w[d], w[b] = rotl(w[b] ^ w[c], 7);
        };
        for (int r = 0; r < rounds / 2; r++) {
            qr(0, 4, 8, 12), qr(1, 5, 9, 13), qr(2, 6, 10, 14), qr(3, 7, 11, 15);
            qr(0, 5, 10, 15), qr(1, 6, 11, 12), qr(2, 7, 8, 13), qr(3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) out[i] = w[i] + in[i];
    }

  private:
    static constexpr size_t kBuf = 64;  // 4 blocks
    void generate() {
        for (size_t b = 0; b < kBuf / 16; b++) {
            uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
            for (int i = 0; i < 8; i++) st[4 + i] = key_[i];
            st[12] = (uint32_t)counter_, st[13] = (uint32_t)(counter_ >> 32), st[14] = (uint32_t)stream_, st[15] = (uint32_t)(stream_ >> 32);
            block(st, rounds_, buf_ + 16 * b);
            counter_++;
        }
    }
    int rounds_;
    uint64_t stream_, counter_ = 0;
    uint32_t key_[8];
    uint32_t buf_[kBuf] = {};
    size_t index_ = kBuf;
};
using StdRng = ChaChaRng;
// ark_std::test_rng() (ark-std 0.4)
inline StdRng test_rng() {
    std::array<uint8_t, 32> seed{};
    const uint8_t head[16] = {1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0};
    std::copy(head, head + 16, seed.begin());
    return StdRng::from_seed(seed);
}

// ---- R1CSToQAP selectors ---------------------------------------------------------------------------------------------------------
struct LibsnarkReduction {  // forks/groth16/src/r1cs_to_qap.rs:100-226; the default of Groth16<E, QAP> and what Crescent uses
    static constexpr int ID = G16_REDUCTION_LIBSNARK;
};
struct CircomReduction {  // forks/circom-compat/src/circom/qap.rs:15-108
    static constexpr int ID = G16_REDUCTION_CIRCOM;
};

// ---- device context (RAII over g16_ctx) ------------------------------------------------------------------------------------------------
class Context {
  public:
    explicit Context(int device = 0, void* stream = nullptr) {
        int rc = g16_ctx_create(&ctx_, device, stream);
        if (rc != G16_OK) throw Panic(rc, std::string("g16_ctx_create: ") + g16_last_error(nullptr));
    }
    ~Context() {
        if (ctx_) g16_ctx_destroy(ctx_);
    }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    g16_ctx* get() const { return ctx_; }
    // maps the C status to the reference's error behaviour
    void check(int rc, const char* what) const {
        if (rc == G16_OK) return;
        std::string msg = std::string(what) + ": " + g16_last_error(ctx_);
        if (rc == G16_ERR_DEGREE_TOO_LARGE) throw SynthesisError(SynthesisError::PolynomialDegreeTooLarge, msg);
        throw Panic(rc, msg);
    }

  private:
    g16_ctx* ctx_ = nullptr;
};

// A full assignment in page-locked memory (g16_host_alloc): what the witness calculator should write into, so that the per-proof
// upload is one DMA (46 MB in 0.7 ms for rs256) instead of a driver-staged copy from pageable memory (~6 ms).
class PinnedAssignment {
  public:
    PinnedAssignment(Context& ctx, size_t len) : ctx_(&ctx), len_(len) {
        void* p = nullptr;
        ctx.check(g16_host_alloc(ctx.get(), len * sizeof(Fr), &p), "g16_host_alloc");
        data_ = static_cast<Fr*>(p);
    }
    ~PinnedAssignment() {
        if (data_) g16_host_free(ctx_->get(), data_);
    }
    PinnedAssignment(const PinnedAssignment&) = delete;
    PinnedAssignment& operator=(const PinnedAssignment&) = delete;
    Fr* data() { return data_; }
    const Fr* data() const { return data_; }
    size_t size() const { return len_; }

  private:
    Context* ctx_;
    Fr* data_ = nullptr;
    size_t len_;
};

// ---- Groth16<Bn254, QAP> prover bound to one GPU ---------------------------------------------------------------------------------------------
// The reference keeps nothing between calls; here an opaque persistent context owns the device copies of the proving key
// queries and of the CSR matrices, keyed by object identity, loaded lazily on first use (SURVEY 8b "Ownership").
template <class QAP = LibsnarkReduction>
class Groth16 {
  public:
    explicit Groth16(int device = 0, bool precompute = true, void* stream = nullptr) : ctx_(device, stream), precompute_(precompute) {}

    // R1CSToQAP::witness_map_from_matrices: the n coefficients of h (Libsnark) / evaluations (Circom), Montgomery
    std::vector<Fr> witness_map_from_matrices(const ConstraintMatrices& matrices, size_t num_inputs, size_t num_constraints,
                                              const std::vector<Fr>& full_assignment) {
        std::lock_guard<std::mutex> g(mu_);
        ensure_matrices(matrices, num_inputs, num_constraints);
        check_assignment(matrices, full_assignment.size());
        size_t n = 0;
        ctx_.check(g16_domain_size(ctx_.get(), &n), "g16_domain_size");
        std::vector<Fr> h(n);
        size_t got = 0;
        ctx_.check(g16_witness_map(ctx_.get(), words(full_assignment), QAP::ID, reinterpret_cast<uint64_t*>(h.data()), n, &got), "g16_witness_map");
        return h;
    }

    // prover.rs:26-51
    Proof create_proof_with_reduction_and_matrices(const ProvingKey& pk, const Fr& r, const Fr& s, const ConstraintMatrices& matrices,
                                                   size_t num_inputs, size_t num_constraints, const std::vector<Fr>& full_assignment) {
        std::lock_guard<std::mutex> g(mu_);
        ensure_matrices(matrices, num_inputs, num_constraints);
        ensure_pk(pk);
        check_assignment(matrices, full_assignment.size());
        g16_proof out{};
        ctx_.check(g16_prove(ctx_.get(), words(full_assignment), r.v.data(), s.v.data(), QAP::ID, &out), "g16_prove");
        return Proof::from_abi(out);
    }

    // same call with the assignment in caller-owned memory (e.g. a PinnedAssignment)
    Proof create_proof_with_reduction_and_matrices(const ProvingKey& pk, const Fr& r, const Fr& s, const ConstraintMatrices& matrices,
                                                   size_t num_inputs, size_t num_constraints, const Fr* full_assignment, size_t len) {
        std::lock_guard<std::mutex> g(mu_);
        ensure_matrices(matrices, num_inputs, num_constraints);
        ensure_pk(pk);
        check_assignment(matrices, len);
        g16_proof out{};
        ctx_.check(g16_prove(ctx_.get(), reinterpret_cast<const uint64_t*>(full_assignment), r.v.data(), s.v.data(), QAP::ID, &out), "g16_prove");
        return Proof::from_abi(out);
    }
    // the matrices of a circuit's R1CS, built once and cached
    const ConstraintMatrices& matrices_of(const std::shared_ptr<const R1CS>& r1cs) { return matrices_for(r1cs); }

    // prover.rs:177-221: the circuit owns the R1CS and the witness.  The matrices are proof-independent, so they are built
    // from the circuit's R1CS once and cached instead of re-synthesised per proof.
    Proof create_proof_with_reduction(const CircomCircuit& circuit, const ProvingKey& pk, const Fr& r, const Fr& s) {
        if (!circuit.witness) throw SynthesisError(SynthesisError::AssignmentMissing, "circuit has no witness");  // circuit.rs:37-42
        const ConstraintMatrices& m = matrices_for(circuit.r1cs);
        return create_proof_with_reduction_and_matrices(pk, r, s, m, m.num_instance_variables, m.num_constraints, *circuit.witness);
    }
    // prover.rs:142-154: r then s drawn from the caller's rng
    template <class Rng>
    Proof create_random_proof_with_reduction(const CircomCircuit& circuit, const ProvingKey& pk, Rng& rng) {
        Fr r = Fr::rand(rng);
        Fr s = Fr::rand(rng);
        return create_proof_with_reduction(circuit, pk, r, s);
    }
    // SNARK::prove (forks/groth16/src/lib.rs:76-82)
    template <class Rng>
    Proof prove(const ProvingKey& pk, const CircomCircuit& circuit, Rng& rng) { return create_random_proof_with_reduction(circuit, pk, rng); }
    // prover.rs:159-172 (r = s = 0; the B-in-G1 MSM is skipped as at prover.rs:102)
    Proof create_proof_with_reduction_no_zk(const CircomCircuit& circuit, const ProvingKey& pk) {
        return create_proof_with_reduction(circuit, pk, Fr::zero(), Fr::zero());
    }

    g16_timings timings() {
        g16_timings t{};
        ctx_.check(g16_get_timings(ctx_.get(), &t), "g16_get_timings");
        return t;
    }
    void set_option(const char* key, int value) { ctx_.check(g16_set_option(ctx_.get(), key, value), "g16_set_option"); }
    uint64_t launch_count() const { return g16_launch_count(ctx_.get()); }
    Context& context() { return ctx_; }

  private:
    static const uint64_t* words(const std::vector<Fr>& v) {
        static_assert(sizeof(Fr) == 32, "Fr must be 4 packed u64 limbs");
        return reinterpret_cast<const uint64_t*>(v.data());
    }
    void check_assignment(const ConstraintMatrices& m, size_t len) {
        if (len != m.num_instance_variables + m.num_witness_variables) throw Panic(G16_ERR_BAD_ARG, "full_assignment length != number of wires");
    }
    void ensure_matrices(const ConstraintMatrices& m, size_t num_inputs, size_t num_constraints) {
        if (num_inputs != m.num_instance_variables || num_constraints != m.num_constraints)
            throw Panic(G16_ERR_BAD_ARG, "num_inputs / num_constraints disagree with the matrices");
        if (loaded_matrices_ == &m) return;
        g16_r1cs_view v = m.view();
        loaded_matrices_ = nullptr;
        ctx_.check(g16_ctx_load_r1cs(ctx_.get(), &v), "g16_ctx_load_r1cs");
        loaded_matrices_ = &m;
    }
    void ensure_pk(const ProvingKey& pk) {
        if (loaded_pk_ == &pk) return;
        g16_pk_view v = pk.view();
        loaded_pk_ = nullptr;
        ctx_.check(g16_ctx_load_pk(ctx_.get(), &v, 0, 1, precompute_ ? 1 : 0), "g16_ctx_load_pk");
        loaded_pk_ = &pk;
    }
    const ConstraintMatrices& matrices_for(const std::shared_ptr<const R1CS>& r1cs) {
        std::lock_guard<std::mutex> g(mu_);
        if (!r1cs) throw Panic(G16_ERR_BAD_ARG, "circuit has no r1cs");
        if (cached_r1cs_ != r1cs) {
            cached_matrices_ = std::make_unique<ConstraintMatrices>(r1cs->to_matrices());
            cached_r1cs_ = r1cs;
            loaded_matrices_ = nullptr;
        }
        return *cached_matrices_;
    }
    Context ctx_;
    bool precompute_;
    std::mutex mu_;
    const ProvingKey* loaded_pk_ = nullptr;
    const ConstraintMatrices* loaded_matrices_ = nullptr;
    std::shared_ptr<const R1CS> cached_r1cs_;
    std::unique_ptr<ConstraintMatrices> cached_matrices_;
};

// ---- the verifier (forks/groth16/src/verifier.rs; scope row f-4) ------------------------------------------------------------------------
// PreparedVerifyingKey (data_structures.rs:62-72): gamma_g2_neg_pc and delta_g2_neg_pc are line-coefficient tables that live
// on the device; the host keeps alpha_g1_beta_g2 (an Fq12 as 12 Montgomery Fq in ark-serialize order) and the key it came from.
struct PreparedVerifyingKey {
    VerifyingKey vk;
    int encoding = G16_ENC_MONTGOMERY;
    std::array<Fq, 12> alpha_g1_beta_g2{};
    size_t num_public_inputs() const { return vk.gamma_abc_g1.size() - 1; }
    std::vector<uint8_t> alpha_g1_beta_g2_bytes() const {  // CanonicalSerialize of the Fq12: 12 x 32 canonical LE bytes
        std::vector<uint8_t> out;
        for (const Fq& f : alpha_g1_beta_g2) detail::put_le(out, f.into_bigint());
        return out;
    }
};

// Verification bound to one GPU.  One call of verify_proofs checks n (proof, public inputs) pairs, one device thread per
// proof; every verdict is what the reference's verify_proof returns for that pair.
class Groth16Verifier {
  public:
    explicit Groth16Verifier(int device = 0, void* stream = nullptr) : ctx_(device, stream) {}

    // verifier.rs:13-20.  `encoding` describes vk's words (G16_ENC_CANONICAL for a key read from arkworks bytes).
    PreparedVerifyingKey prepare_verifying_key(const VerifyingKey& vk, int encoding) {
        std::lock_guard<std::mutex> g(mu_);
        PreparedVerifyingKey pvk;
        pvk.vk = vk;
        pvk.encoding = encoding;
        load(pvk);
        uint64_t gt[48];
        ctx_.check(g16_vk_alpha_beta(ctx_.get(), gt), "g16_vk_alpha_beta");
        for (int i = 0; i < 12; i++) pvk.alpha_g1_beta_g2[i] = Fq{Limbs{gt[4 * i], gt[4 * i + 1], gt[4 * i + 2], gt[4 * i + 3]}};
        return pvk;
    }
    PreparedVerifyingKey prepare_verifying_key(const ProvingKey& pk) { return prepare_verifying_key(pk.vk, pk.encoding); }

    // verifier.rs:25-39
    G1Affine prepare_inputs(const PreparedVerifyingKey& pvk, const std::vector<Fr>& public_inputs) {
        if (public_inputs.size() + 1 != pvk.vk.gamma_abc_g1.size())
            throw SynthesisError(SynthesisError::MalformedVerifyingKey, "public input count != gamma_abc_g1.len() - 1");
        std::lock_guard<std::mutex> g(mu_);
        ensure(pvk);
        uint64_t out[8];
        ctx_.check(g16_prepare_inputs(ctx_.get(), reinterpret_cast<const uint64_t*>(public_inputs.data()), 1, out), "g16_prepare_inputs");
        bool inf = true;
        for (uint64_t w : out) inf = inf && w == 0;
        if (inf) return G1Affine::identity();
        return G1Affine{Fq{Limbs{out[0], out[1], out[2], out[3]}}, Fq{Limbs{out[4], out[5], out[6], out[7]}}, false};
    }

    // verifier.rs:44-65
    bool verify_proof_with_prepared_inputs(const PreparedVerifyingKey& pvk, const Proof& proof, const G1Affine& prepared_inputs) {
        std::lock_guard<std::mutex> g(mu_);
        ensure(pvk);
        g16_proof p = proof.to_abi();
        uint64_t pi[8] = {};
        if (!prepared_inputs.infinity) {
            std::memcpy(pi, prepared_inputs.x.v.data(), 32);
            std::memcpy(pi + 4, prepared_inputs.y.v.data(), 32);
        }
        uint8_t verdict = 0;
        ctx_.check(g16_verify_batch_prepared(ctx_.get(), &p, pi, 1, &verdict), "g16_verify_batch_prepared");
        return decide(verdict);
    }

    // verifier.rs:69-76
    bool verify_proof(const PreparedVerifyingKey& pvk, const Proof& proof, const std::vector<Fr>& public_inputs) {
        return verify_proofs(pvk, {proof}, {public_inputs})[0];
    }

    // SNARK trait names (forks/groth16/src/lib.rs:84-96)
    PreparedVerifyingKey process_vk(const VerifyingKey& circuit_vk, int encoding) { return prepare_verifying_key(circuit_vk, encoding); }
    bool verify_with_processed_vk(const PreparedVerifyingKey& circuit_pvk, const std::vector<Fr>& x, const Proof& proof) {
        return verify_proof(circuit_pvk, proof, x);
    }

    // n independent verify_proof calls in one launch
    std::vector<bool> verify_proofs(const PreparedVerifyingKey& pvk, const std::vector<Proof>& proofs,
                                    const std::vector<std::vector<Fr>>& public_inputs) {
        if (proofs.size() != public_inputs.size()) throw Panic(G16_ERR_BAD_ARG, "one public-input vector per proof");
        const size_t k = pvk.num_public_inputs();
        std::vector<g16_proof> ps;
        std::vector<uint64_t> xs;
        ps.reserve(proofs.size());
        xs.reserve(proofs.size() * k * 4);
        for (size_t i = 0; i < proofs.size(); i++) {
            if (public_inputs[i].size() != k)
                throw SynthesisError(SynthesisError::MalformedVerifyingKey, "public input count != gamma_abc_g1.len() - 1");
            ps.push_back(proofs[i].to_abi());
            for (const Fr& x : public_inputs[i]) xs.insert(xs.end(), x.v.begin(), x.v.end());
        }
        std::vector<bool> out(proofs.size());
        if (proofs.empty()) return out;
        std::vector<uint8_t> verdict(proofs.size());
        {
            std::lock_guard<std::mutex> g(mu_);
            ensure(pvk);
            ctx_.check(g16_verify_batch(ctx_.get(), ps.data(), k ? xs.data() : nullptr, ps.size(), verdict.data()), "g16_verify_batch");
        }
        for (size_t i = 0; i < verdict.size(); i++) out[i] = decide(verdict[i]);
        return out;
    }
    uint64_t launch_count() const { return g16_launch_count(ctx_.get()); }
    Context& context() { return ctx_; }

  private:
    static bool decide(uint8_t verdict) {
        if (verdict == G16_VERDICT_UNEXPECTED_IDENTITY) throw SynthesisError(SynthesisError::UnexpectedIdentity, "final exponentiation of zero");
        return verdict == G16_VERDICT_ACCEPT;
    }
    static std::vector<uint64_t> key_words(const PreparedVerifyingKey& pvk) {
        std::vector<uint64_t> k{(uint64_t)pvk.encoding};
        for (const PointVec* p : {&pvk.vk.alpha_g1, &pvk.vk.beta_g2, &pvk.vk.gamma_g2, &pvk.vk.delta_g2, &pvk.vk.gamma_abc_g1})
            k.insert(k.end(), p->w.begin(), p->w.end());
        return k;
    }
    void load(const PreparedVerifyingKey& pvk) {
        if (pvk.vk.alpha_g1.size() != 1 || pvk.vk.beta_g2.size() != 1 || pvk.vk.gamma_g2.size() != 1 || pvk.vk.delta_g2.size() != 1 ||
            pvk.vk.gamma_abc_g1.size() == 0)
            throw SynthesisError(SynthesisError::MalformedVerifyingKey, "verifying key is missing a point");
        g16_vk_view v{};
        v.alpha_g1 = pvk.vk.alpha_g1.data(), v.beta_g2 = pvk.vk.beta_g2.data(), v.gamma_g2 = pvk.vk.gamma_g2.data();
        v.delta_g2 = pvk.vk.delta_g2.data(), v.gamma_abc_g1 = pvk.vk.gamma_abc_g1.data(), v.gamma_abc_len = pvk.vk.gamma_abc_g1.size();
        v.encoding = pvk.encoding;
        ctx_.check(g16_ctx_load_vk(ctx_.get(), &v), "g16_ctx_load_vk");
        loaded_ = key_words(pvk);
    }
    // the device holds one prepared key: a pvk other than the one loaded last is prepared again on first use
    void ensure(const PreparedVerifyingKey& pvk) {
        if (loaded_.empty() || loaded_ != key_words(pvk)) load(pvk);
    }
    Context ctx_;
    std::mutex mu_;
    std::vector<uint64_t> loaded_;
};

// ---- MSM-sharded proving inside one process: one context (and one host thread per call) per shard ------------------------------------------
// Shard k keeps the point range [k*N/G, (k+1)*N/G) of each query (g16_ctx_load_pk), runs the witness map and its five
// partial MSMs (g16_prove_shard); the G 896-byte partials are added on shard 0's device (g16_prove_combine).  `devices`
// may name the same GPU several times (the shards then share it) -- that is how the logic is tested on a one-GPU box.
// The one-process-per-GPU variant with an NCCL gather is crescent_credentials_b200/sharded.py.
template <class QAP = LibsnarkReduction>
class ShardedGroth16 {
  public:
    ShardedGroth16(const std::vector<int>& devices, const ProvingKey& pk, const ConstraintMatrices& matrices, bool precompute = true)
        : wires_(matrices.num_instance_variables + matrices.num_witness_variables) {
        if (devices.empty()) throw Panic(G16_ERR_BAD_ARG, "no devices");
        g16_pk_view pv = pk.view();
        g16_r1cs_view rv = matrices.view();
        for (size_t k = 0; k < devices.size(); k++) {
            ctxs_.push_back(std::make_unique<Context>(devices[k]));
            ctxs_[k]->check(g16_ctx_load_r1cs(ctxs_[k]->get(), &rv), "g16_ctx_load_r1cs");
            ctxs_[k]->check(g16_ctx_load_pk(ctxs_[k]->get(), &pv, (int)k, (int)devices.size(), precompute ? 1 : 0), "g16_ctx_load_pk");
        }
    }
    Proof prove(const Fr& r, const Fr& s, const std::vector<Fr>& full_assignment) {
        if (full_assignment.size() != wires_) throw Panic(G16_ERR_BAD_ARG, "full_assignment length != number of wires");
        const uint64_t* z = reinterpret_cast<const uint64_t*>(full_assignment.data());
        size_t G = ctxs_.size();
        std::vector<g16_partial> partials(G);
        std::vector<int> rc(G, G16_OK);
        ctxs_[0]->check(g16_prove_prepare(ctxs_[0]->get(), r.v.data(), s.v.data()), "g16_prove_prepare");
        std::vector<std::thread> th;
        for (size_t k = 1; k < G; k++)
            th.emplace_back([&, k] { rc[k] = g16_prove_shard(ctxs_[k]->get(), z, r.v.data(), s.v.data(), QAP::ID, &partials[k]); });
        rc[0] = g16_prove_shard(ctxs_[0]->get(), z, r.v.data(), s.v.data(), QAP::ID, &partials[0]);
        for (auto& t : th) t.join();
        for (size_t k = 0; k < G; k++) ctxs_[k]->check(rc[k], "g16_prove_shard");
        g16_proof out{};
        ctxs_[0]->check(g16_prove_combine(ctxs_[0]->get(), partials.data(), (int)G, r.v.data(), s.v.data(), &out), "g16_prove_combine");
        return Proof::from_abi(out);
    }

  private:
    size_t wires_;
    std::vector<std::unique_ptr<Context>> ctxs_;
};

// Bulk canonical -> Montgomery conversion of a witness (m elements) on the GPU (g16_field_op): the wasm witness calculator
// hands over canonical integers, `Fr::from` in the reference does this conversion on the CPU.
inline std::vector<Fr> fr_from_canonical_bulk(Context& ctx, const uint8_t* le_bytes, size_t count) {
    std::vector<Fr> out(count);
    if (!count) return out;
    std::vector<uint64_t> in(count * 4);
    std::memcpy(in.data(), le_bytes, count * 32);
    ctx.check(g16_field_op(ctx.get(), G16_FIELD_FR, G16_OP_TO_MONT, in.data(), nullptr, reinterpret_cast<uint64_t*>(out.data()), count),
              "g16_field_op(to_mont)");
    return out;
}

}  // namespace ark_groth16_b200
#endif  // ARK_GROTH16_B200_HPP
