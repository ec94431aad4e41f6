// g16_cli -- small driver over ark_groth16_b200.hpp: the compiled host side of the drop-in, used by the tests to run the
// reference-shaped calls (CircomCircuit + ProvingKey bytes in, ark-serialize Proof bytes out) without Python in between.
// It mirrors what `crescent prove` does around Groth16::prove (creds/src/lib.rs:255-303): read main_c.r1cs, read
// prover_params.bin, take the witness, prove, write the proof -- nothing else of the CLI (no JWT handling, no show/verify).
//
//   g16_cli prove --r1cs F --pk F --witness F (--r HEX --s HEX | --seed N | --test-rng | --no-zk) [--reduction circom]
//                 [--no-precompute] [--shards K] [--device D] [--repeat N] [--pageable] --out F
//   g16_cli witness-map --r1cs F --witness F [--reduction circom] --out F      (n x 32 bytes, canonical little-endian)
//   g16_cli rng (test | seed:N) COUNT          next_u64 draws of StdRng, one hex word per line
//   g16_cli rand-fr (test | seed:N) COUNT      Fr::rand draws, canonical value as hex
//   g16_cli r1cs F                              to_matrices(): "k row col coeff" lines
//   g16_cli pk-roundtrip IN OUT                 deserialize_uncompressed_unchecked -> serialize_uncompressed
//   g16_cli proof-ser AX AY BX0 BX1 BY0 BY1 CX CY    canonical coordinates ("inf" for a point: AX=inf AY=-)
//   g16_cli fp (fr|fq) (from|into) HEX          host Montgomery marshalling check
//   g16_cli r1cs-write --prefix P --nc N --nwires M --ninputs L --out F    CSR dumps (P.{0,1,2}.{ptr,col,val}) -> iden3 .r1cs
//   g16_cli pk-write --prefix P --out F          canonical point dumps (P.<field>.bin) -> arkworks uncompressed ProvingKey bytes
//   g16_cli r1cs-bench F                         R1CSFile::read + R1CS::to_matrices timings (scope row f-1)
// r1cs-write / pk-write exist so that full-size fixtures in the reference's FILE formats can be produced from the synthetic
// instance of bench.py (tools/next_rows_bench.py): the prove command then runs the drop-in flow end to end on them.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "ark_groth16_b200.hpp"

using namespace ark_groth16_b200;

static std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<uint8_t> b((size_t)f.tellg());  // one sized read: the key and the .r1cs file are ~0.6 GB each
    f.seekg(0);
    if (!b.empty() && !f.read(reinterpret_cast<char*>(b.data()), (std::streamsize)b.size())) throw std::runtime_error("cannot read " + path);
    return b;
}
static void write_file(const std::string& path, const std::vector<uint8_t>& b) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + path);
    f.write(reinterpret_cast<const char*>(b.data()), (std::streamsize)b.size());
}
static std::string hex_bytes(const std::vector<uint8_t>& b) {
    static const char* d = "0123456789abcdef";
    std::string s;
    for (uint8_t x : b) s.push_back(d[x >> 4]), s.push_back(d[x & 15]);
    return s;
}
static StdRng make_rng(const std::string& spec) {
    if (spec == "test") return test_rng();
    if (spec.rfind("seed:", 0) == 0) return StdRng::seed_from_u64(std::stoull(spec.substr(5), nullptr, 0));
    throw std::invalid_argument("rng spec must be `test` or `seed:N`");
}

struct Args {
    std::map<std::string, std::string> kv;
    std::vector<std::string> pos;
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    std::string get(const std::string& k, const std::string& dflt = "") const { return has(k) ? kv.at(k) : dflt; }
    std::string req(const std::string& k) const {
        if (!has(k)) throw std::invalid_argument("missing --" + k);
        return kv.at(k);
    }
};
static Args parse(int argc, char** argv, int from) {
    static const char* flags[] = {"test-rng", "no-zk", "no-precompute", "pageable"};
    Args a;
    for (int i = from; i < argc; i++) {
        std::string s = argv[i];
        if (s.rfind("--", 0) == 0) {
            std::string k = s.substr(2);
            bool is_flag = false;
            for (auto f : flags) is_flag = is_flag || k == f;
            if (is_flag) a.kv[k] = "1";
            else if (i + 1 < argc) a.kv[k] = argv[++i];
            else throw std::invalid_argument("missing value for " + s);
        } else
            a.pos.push_back(s);
    }
    return a;
}

static Fr fr_from_hex(const std::string& h) {
    auto v = Fr::from_bigint(detail::from_hex(h));
    if (!v) throw std::invalid_argument("scalar not below the Fr modulus");
    return *v;
}

static double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
template <class T>
static std::vector<T> read_raw(const std::string& path) {
    std::vector<uint8_t> b = read_file(path);
    if (b.size() % sizeof(T)) throw std::runtime_error(path + ": size is not a multiple of the element size");
    std::vector<T> v(b.size() / sizeof(T));
    std::memcpy(v.data(), b.data(), b.size());
    return v;
}

// CSR (canonical coefficient words) -> iden3 .r1cs v1 in the layout r1cs_reader.rs:54-256 parses
static int run_r1cs_write(const Args& a) {
    const std::string prefix = a.req("prefix");
    const uint32_t nc = (uint32_t)std::stoull(a.req("nc")), nwires = (uint32_t)std::stoull(a.req("nwires"));
    const uint32_t ninputs = (uint32_t)std::stoull(a.req("ninputs"));
    std::vector<uint64_t> ptr[3], val[3];
    std::vector<uint32_t> col[3];
    size_t cons_bytes = 0;
    for (int k = 0; k < 3; k++) {
        ptr[k] = read_raw<uint64_t>(prefix + "." + std::to_string(k) + ".ptr");
        col[k] = read_raw<uint32_t>(prefix + "." + std::to_string(k) + ".col");
        val[k] = read_raw<uint64_t>(prefix + "." + std::to_string(k) + ".val");
        if (ptr[k].size() != (size_t)nc + 1 || ptr[k][nc] != col[k].size() || val[k].size() != col[k].size() * 4)
            throw std::runtime_error("CSR dump " + std::to_string(k) + " is inconsistent");
        cons_bytes += (size_t)nc * 4 + col[k].size() * 36;
    }
    std::vector<uint8_t> out;
    out.reserve(12 + 3 * 12 + 64 + cons_bytes + (size_t)nwires * 8);
    auto u32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) out.push_back((uint8_t)(v >> (8 * i))); };
    auto u64 = [&](uint64_t v) { for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i))); };
    out.insert(out.end(), {'r', '1', 'c', 's'});
    u32(1), u32(3);
    u32(1), u64(64);  // header section
    u32(32);
    detail::put_le(out, detail::kFrModulus);
    u32(nwires), u32(ninputs - 1) /* n_pub_out */, u32(0) /* n_pub_in */, u32(nwires - ninputs) /* n_prv_in */, u64(nwires), u32(nc);
    u32(2), u64(cons_bytes);  // constraints
    for (uint32_t i = 0; i < nc; i++)
        for (int k = 0; k < 3; k++) {
            u32((uint32_t)(ptr[k][i + 1] - ptr[k][i]));
            for (uint64_t j = ptr[k][i]; j < ptr[k][i + 1]; j++) {
                u32(col[k][j]);
                for (int w = 0; w < 4; w++) u64(val[k][4 * j + w]);
            }
        }
    u32(3), u64((uint64_t)nwires * 8);  // wire -> label map (identity)
    for (uint32_t i = 0; i < nwires; i++) u64(i);
    write_file(a.req("out"), out);
    std::cout << out.size() << "\n";
    return 0;
}

// canonical point dumps -> arkworks' uncompressed ProvingKey bytes (data_structures.rs:31-44,101-118 field order; flags as
// ark-serialize writes them: bit 6 of a point's last byte = infinity, bit 7 = "y is the larger of {y, -y}")
static int run_pk_write(const Args& a) {
    const std::string prefix = a.req("prefix");
    std::vector<uint8_t> out;
    auto points = [&](const char* name, int words, bool with_len) {
        std::vector<uint64_t> w = read_raw<uint64_t>(prefix + "." + name + ".bin");
        if (w.size() % words) throw std::runtime_error(std::string(name) + ": not a whole number of points");
        size_t n = w.size() / words;
        if (!with_len && n != 1) throw std::runtime_error(std::string(name) + ": expected one point");
        if (with_len)
            for (int k = 0; k < 8; k++) out.push_back((uint8_t)((uint64_t)n >> (8 * k)));
        size_t at = out.size();
        out.resize(at + w.size() * 8);
        std::memcpy(out.data() + at, w.data(), w.size() * 8);
        for (size_t i = 0; i < n; i++) {
            const uint64_t* p = w.data() + i * words;
            bool inf = true;
            for (int k = 0; k < words; k++) inf = inf && p[k] == 0;
            uint8_t flag;
            if (inf) flag = detail::kFlagInfinity;
            else if (words == 8) flag = detail::fq_is_negative(Limbs{p[4], p[5], p[6], p[7]}) ? detail::kFlagNegative : 0;
            else flag = detail::fq2_is_negative(Limbs{p[8], p[9], p[10], p[11]}, Limbs{p[12], p[13], p[14], p[15]}) ? detail::kFlagNegative : 0;
            out[at + (i + 1) * words * 8 - 1] |= flag;
        }
    };
    points("alpha_g1", 8, false), points("beta_g2", 16, false), points("gamma_g2", 16, false), points("vk_delta_g1", 8, false);
    points("delta_g2", 16, false), points("gamma_abc_g1", 8, true), points("beta_g1", 8, false), points("delta_g1", 8, false);
    points("a_query", 8, true), points("b_g1_query", 8, true), points("b_g2_query", 16, true), points("h_query", 8, true);
    points("l_query", 8, true);
    write_file(a.req("out"), out);
    std::cout << out.size() << "\n";
    return 0;
}

// scope row f-1: the matrices are proof-independent, so the file is parsed and flattened ONCE per circuit
static int run_r1cs_bench(const Args& a) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<uint8_t> bytes = read_file(a.pos.at(0));
    double t_file = ms_since(t0);
    t0 = std::chrono::steady_clock::now();
    R1CS r1cs = R1CS::from_file(R1CSFile::read(bytes));
    double t_parse = ms_since(t0);
    t0 = std::chrono::steady_clock::now();
    ConstraintMatrices m = r1cs.to_matrices();
    double t_mat = ms_since(t0);
    printf("{\"file_bytes\": %zu, \"constraints\": %zu, \"nnz\": [%zu, %zu, %zu], \"read_file_ms\": %.1f, \"parse_ms\": %.1f, "
           "\"to_matrices_ms\": %.1f}\n", bytes.size(), m.num_constraints, m.a_num_non_zero(), m.b_num_non_zero(), m.c_num_non_zero(),
           t_file, t_parse, t_mat);
    return 0;
}

template <class QAP>
static int run_prove(const Args& a) {
    auto t_start = std::chrono::steady_clock::now();
    auto r1cs = std::make_shared<const R1CS>(R1CS::from_file(R1CSFile::read(read_file(a.req("r1cs")))));
    double t_r1cs = ms_since(t_start);
    t_start = std::chrono::steady_clock::now();
    ProvingKey pk = ProvingKey::deserialize_uncompressed_unchecked(read_file(a.req("pk")));
    double t_pk = ms_since(t_start);
    std::vector<uint8_t> wbytes = read_file(a.req("witness"));
    if (wbytes.size() != r1cs->num_variables * 32) throw std::invalid_argument("witness file must hold num_wires x 32 bytes");
    int device = std::stoi(a.get("device", "0"));
    bool precompute = !a.has("no-precompute");
    int repeat = std::stoi(a.get("repeat", "1"));
    int shards = std::stoi(a.get("shards", "1"));

    Fr r = Fr::zero(), s = Fr::zero();
    if (a.has("seed") || a.has("test-rng")) {
        StdRng rng = a.has("test-rng") ? test_rng() : StdRng::seed_from_u64(std::stoull(a.req("seed"), nullptr, 0));
        r = Fr::rand(rng);  // prover.rs:151-152: r first, then s
        s = Fr::rand(rng);
    } else if (!a.has("no-zk")) {
        r = fr_from_hex(a.req("r"));
        s = fr_from_hex(a.req("s"));
    }

    Proof proof;
    double ms = 0, first_ms = 0;
    if (shards <= 1) {
        Groth16<QAP> prover(device, precompute);
        CircomCircuit circuit{r1cs, fr_from_canonical_bulk(prover.context(), wbytes.data(), r1cs->num_variables)};
        // by default the assignment lives in page-locked memory (what a witness calculator bound to this library would fill);
        // --pageable proves straight from the std::vector of the CircomCircuit
        std::unique_ptr<PinnedAssignment> pinned;
        if (!a.has("pageable")) {
            pinned = std::make_unique<PinnedAssignment>(prover.context(), circuit.witness->size());
            std::memcpy(pinned->data(), circuit.witness->data(), circuit.witness->size() * sizeof(Fr));
        }
        for (int k = 0; k < repeat; k++) {
            auto t0 = std::chrono::steady_clock::now();
            Proof p;
            if (pinned) {
                const ConstraintMatrices& m = prover.matrices_of(circuit.r1cs);
                p = prover.create_proof_with_reduction_and_matrices(pk, r, s, m, m.num_instance_variables, m.num_constraints, pinned->data(),
                                                                    pinned->size());
            } else
                p = prover.create_proof_with_reduction(circuit, pk, r, s);
            ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (k == 0) first_ms = ms;  // includes to_matrices, the CSR / key uploads and the window tables
            if (k && !(p == proof)) throw std::runtime_error("proof changed between repeats");
            proof = p;
        }
        fprintf(stderr, "g16_cli: %llu kernel launches, last prove %.3f ms (host clock)\n", (unsigned long long)prover.launch_count(), ms);
        fprintf(stderr, "g16_cli timings: {\"r1cs_read_parse_ms\": %.1f, \"pk_read_deserialize_ms\": %.1f, \"first_prove_ms\": %.1f, "
                        "\"last_prove_ms\": %.3f}\n", t_r1cs, t_pk, first_ms, ms);
    } else {
        ConstraintMatrices m = r1cs->to_matrices();
        std::vector<int> devices;
        int ndev = g16_device_count();
        for (int k = 0; k < shards; k++) devices.push_back(ndev > 0 ? (device + k) % ndev : device + k);
        Context conv(device);
        std::vector<Fr> z = fr_from_canonical_bulk(conv, wbytes.data(), r1cs->num_variables);
        ShardedGroth16<QAP> prover(devices, pk, m, precompute);
        for (int k = 0; k < repeat; k++) {
            Proof p = prover.prove(r, s, z);
            if (k && !(p == proof)) throw std::runtime_error("proof changed between repeats");
            proof = p;
        }
    }
    write_file(a.req("out"), proof.serialize_uncompressed());
    std::cout << hex_bytes(proof.serialize_compressed()) << "\n";
    return 0;
}

// verify: arkworks pk bytes (the vk is its prefix), 256-byte uncompressed proofs, public inputs as hex -> one verdict per proof.
//   g16_cli verify --pk key.bin --proof p1.bin[,p2.bin...] --inputs hex,hex,...[;hex,hex,...]  [--gt-out file]
// Prints "1"/"0" per proof on stdout (verify_proof, verifier.rs:69-76) and the hex of alpha_g1_beta_g2 on stderr.
static std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) {
        if (c == sep) {
            out.push_back(cur);
            cur.clear();
        } else
            cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}
static int run_verify(const Args& a) {
    ProvingKey pk = ProvingKey::deserialize_uncompressed_unchecked(read_file(a.req("pk")));
    Groth16Verifier ver(std::stoi(a.get("device", "0")));
    PreparedVerifyingKey pvk = ver.prepare_verifying_key(pk);
    std::vector<Proof> proofs;
    for (const std::string& f : split(a.req("proof"), ',')) {
        std::vector<uint8_t> b = read_file(f);
        proofs.push_back(Proof::deserialize_uncompressed_unchecked(b.data(), b.size()));
    }
    std::vector<std::vector<Fr>> inputs;
    for (const std::string& row : split(a.get("inputs", ""), ';')) {
        std::vector<Fr> x;
        if (!row.empty())
            for (const std::string& h : split(row, ',')) x.push_back(fr_from_hex(h));
        inputs.push_back(x);
    }
    while (inputs.size() < proofs.size()) inputs.push_back(inputs.back());  // one input row for all proofs
    if (a.has("gt-out")) write_file(a.req("gt-out"), pvk.alpha_g1_beta_g2_bytes());
    if (a.has("prepared-out")) {
        std::vector<uint8_t> out;
        serialize_g1(ver.prepare_inputs(pvk, inputs[0]), Compress::No, out);
        write_file(a.req("prepared-out"), out);
    }
    std::vector<bool> ok = ver.verify_proofs(pvk, proofs, inputs);
    // the single-proof entry points must agree with the batch
    if (ver.verify_proof(pvk, proofs[0], inputs[0]) != ok[0] ||
        ver.verify_proof_with_prepared_inputs(pvk, proofs[0], ver.prepare_inputs(pvk, inputs[0])) != ok[0])
        throw std::runtime_error("verify_proof disagrees with verify_proofs");
    for (bool v : ok) std::cout << (v ? "1" : "0") << "\n";
    fprintf(stderr, "g16_cli: %llu kernel launches\n", (unsigned long long)ver.launch_count());
    return 0;
}

template <class QAP>
static int run_witness_map(const Args& a) {
    R1CS r1cs = R1CS::from_file(R1CSFile::read(read_file(a.req("r1cs"))));
    ConstraintMatrices m = r1cs.to_matrices();
    std::vector<uint8_t> wbytes = read_file(a.req("witness"));
    if (wbytes.size() != r1cs.num_variables * 32) throw std::invalid_argument("witness file must hold num_wires x 32 bytes");
    Groth16<QAP> prover(std::stoi(a.get("device", "0")), false);
    std::vector<Fr> z = fr_from_canonical_bulk(prover.context(), wbytes.data(), r1cs.num_variables);
    std::vector<Fr> h = prover.witness_map_from_matrices(m, m.num_instance_variables, m.num_constraints, z);
    std::vector<uint64_t> canon(h.size() * 4);
    Context& ctx = prover.context();
    ctx.check(g16_field_op(ctx.get(), G16_FIELD_FR, G16_OP_FROM_MONT, reinterpret_cast<const uint64_t*>(h.data()), nullptr, canon.data(), h.size()),
              "g16_field_op(from_mont)");
    std::vector<uint8_t> out(canon.size() * 8);
    std::memcpy(out.data(), canon.data(), out.size());
    write_file(a.req("out"), out);
    std::cout << h.size() << "\n";
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: g16_cli (prove|verify|witness-map|rng|rand-fr|r1cs|pk-roundtrip|proof-ser|fp) ...  (see g16_cli.cpp)\n");
        return 64;
    }
    std::string cmd = argv[1];
    try {
        Args a = parse(argc, argv, 2);
        if (cmd == "prove") return a.get("reduction", "libsnark") == "circom" ? run_prove<CircomReduction>(a) : run_prove<LibsnarkReduction>(a);
        if (cmd == "verify") return run_verify(a);
        if (cmd == "r1cs-write") return run_r1cs_write(a);
        if (cmd == "pk-write") return run_pk_write(a);
        if (cmd == "r1cs-bench") return run_r1cs_bench(a);
        if (cmd == "witness-map")
            return a.get("reduction", "libsnark") == "circom" ? run_witness_map<CircomReduction>(a) : run_witness_map<LibsnarkReduction>(a);
        if (cmd == "rng" || cmd == "rand-fr") {
            StdRng rng = make_rng(a.pos.at(0));
            int count = std::stoi(a.pos.at(1));
            for (int i = 0; i < count; i++) {
                if (cmd == "rng") printf("%016llx\n", (unsigned long long)rng.next_u64());
                else std::cout << detail::to_hex(Fr::rand(rng).into_bigint()) << "\n";
            }
            return 0;
        }
        if (cmd == "r1cs") {
            R1CS r = R1CS::from_file(R1CSFile::read(read_file(a.pos.at(0))));
            ConstraintMatrices m = r.to_matrices();
            printf("num_instance %zu num_witness %zu num_constraints %zu nnz %zu %zu %zu\n", m.num_instance_variables, m.num_witness_variables,
                   m.num_constraints, m.a_num_non_zero(), m.b_num_non_zero(), m.c_num_non_zero());
            const Csr* mats[3] = {&m.a, &m.b, &m.c};
            for (int k = 0; k < 3; k++)
                for (size_t i = 0; i < m.num_constraints; i++)
                    for (uint64_t e = mats[k]->row_ptr[i]; e < mats[k]->row_ptr[i + 1]; e++) {
                        Limbs v = {mats[k]->val[4 * e], mats[k]->val[4 * e + 1], mats[k]->val[4 * e + 2], mats[k]->val[4 * e + 3]};
                        printf("%d %zu %u %s\n", k, i, mats[k]->col[e], detail::to_hex(v).c_str());
                    }
            return 0;
        }
        if (cmd == "pk-roundtrip") {
            ProvingKey pk = ProvingKey::deserialize_uncompressed_unchecked(read_file(a.pos.at(0)));
            write_file(a.pos.at(1), pk.serialize_uncompressed());
            printf("a %zu b_g1 %zu b_g2 %zu h %zu l %zu gamma_abc %zu\n", pk.a_query.size(), pk.b_g1_query.size(), pk.b_g2_query.size(),
                   pk.h_query.size(), pk.l_query.size(), pk.vk.gamma_abc_g1.size());
            return 0;
        }
        if (cmd == "proof-ser") {
            auto fq = [&](size_t i) {
                auto v = Fq::from_bigint(detail::from_hex(a.pos.at(i)));
                if (!v) throw std::invalid_argument("coordinate not below the Fq modulus");
                return *v;
            };
            Proof p;
            if (a.pos.at(0) != "inf") p.a = G1Affine{fq(0), fq(1), false};
            if (a.pos.at(2) != "inf") p.b = G2Affine{Fq2{fq(2), fq(3)}, Fq2{fq(4), fq(5)}, false};
            if (a.pos.at(6) != "inf") p.c = G1Affine{fq(6), fq(7), false};
            std::cout << hex_bytes(p.serialize_compressed()) << "\n" << hex_bytes(p.serialize_uncompressed()) << "\n";
            return 0;
        }
        if (cmd == "fp") {
            bool fr = a.pos.at(0) == "fr";
            Limbs x = detail::from_hex(a.pos.at(2));
            Limbs y;
            if (a.pos.at(1) == "from") {
                if (fr) {
                    auto v = Fr::from_bigint(x);
                    if (!v) throw std::invalid_argument("not reduced");
                    y = v->v;
                } else {
                    auto v = Fq::from_bigint(x);
                    if (!v) throw std::invalid_argument("not reduced");
                    y = v->v;
                }
            } else
                y = fr ? Fr{x}.into_bigint() : Fq{x}.into_bigint();
            std::cout << detail::to_hex(y) << "\n";
            return 0;
        }
        fprintf(stderr, "g16_cli: unknown command %s\n", cmd.c_str());
        return 64;
    } catch (const SynthesisError& e) {
        fprintf(stderr, "g16_cli: SynthesisError::%s: %s\n", e.kind == SynthesisError::PolynomialDegreeTooLarge ? "PolynomialDegreeTooLarge" : "AssignmentMissing", e.what());
        return 3;
    } catch (const Panic& e) {
        fprintf(stderr, "g16_cli: panic (code %d): %s\n", e.code, e.what());
        return 100 + e.code;
    } catch (const std::exception& e) {
        fprintf(stderr, "g16_cli: %s\n", e.what());
        return 2;
    }
}
