// oracle/spartan_oracle.cpp -- literal single-threaded CPU restatement of the r1cs-spartan prover.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / `--impl reference` leg may load this library; the product
// (r1cs-spartan_b200/) never links, imports or executes it.
//
// PARITY UNPINNED.  The reference (tsunrise/r1cs-spartan) is Rust over un-pinned arkworks git
// dependencies (/root/reference/Cargo.toml:10-16,22); no cargo/rustc here, no golden vectors in the
// reference.  This file follows the reference's own sources line by line and restates the upstream
// (arkworks ~Dec 2020) behaviour it calls from that project's published algorithms; every such
// piece is tagged UPSTREAM and isolated in one function.  It is cross-checked against an independent
// Python big-integer model (oracle/pymodel.py) and the identities of the reference's own tests.
//
// It deliberately keeps the reference's algorithmic shape (no shortcuts), so it can validate the
// algebraic rewrites the CUDA path uses AND serve as the timed CPU baseline:
//   * eq_extension: log_n separate tables            (src/data_structures/eq.rs:5-20)
//   * first sumcheck: 2 products over 2 log_n + 3 tables, log_n + 3 evaluations per round
//                                                     (src/ahp/prover.rs:163-207)
//   * second sumcheck: 3 products x 2 tables          (src/ahp/prover.rs:230-266)
//   * open: duplicated-scalar G2 MSMs of size 2^k     (src/commitment/open.rs:37-51)
//   * one thread (Cargo.toml:26 never enables `parallel`).
#include "ff.hpp"
#include "ec.hpp"
#include "blake2s.hpp"
#include "pairing.hpp"
#include <vector>
#include <map>
#include <unordered_map>
#include <string>
#include <chrono>
#include <cstdlib>
#include <cassert>

typedef std::vector<std::pair<Fr, size_t>> Row;   // ark_relations::r1cs::Matrix row: (coeff, column)
typedef std::vector<Row> Matrix;

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ deterministic workload PRNG
struct SplitMix64 {
    uint64_t s;
    uint64_t next_u64() {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
};
// UPSTREAM ark_ff `impl Distribution<Fp256<P>> for Standard`: 4 x next_u64 into the limbs, clear the top
// REPR_SHAVE_BITS = 1 bit, retry until < modulus; the accepted limbs ARE the Montgomery residue.
template <typename RNG>
static Fr fr_rand(RNG& rng) {
    for (;;) {
        Fr x;
        for (int i = 0; i < 4; i++) x.l[i] = rng.next_u64();
        x.l[3] &= 0x7fffffffffffffffULL;
        if (!Fr::geq_mod(x.l)) return x;
    }
}

// ------------------------------------------------------------------ UPSTREAM CanonicalSerialize
typedef std::vector<uint8_t> Bytes;
static void ser_u64(Bytes& o, uint64_t v) { for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i))); }
static void ser_fr(Bytes& o, const Fr& x) {
    uint64_t c[4]; x.to_canonical(c);
    for (int i = 0; i < 4; i++) ser_u64(o, c[i]);
}
static void ser_fr_vec(Bytes& o, const std::vector<Fr>& v) { ser_u64(o, v.size()); for (auto& x : v) ser_fr(o, x); }
static void ser_fq_raw(Bytes& o, const Fq& x) {
    uint64_t c[6]; x.to_canonical(c);
    for (int i = 0; i < 6; i++) ser_u64(o, c[i]);
}
// SWFlags: bit 7 = "y > -y" (PositiveY), bit 6 = infinity; OR-ed into the last byte of x.
static void ser_g1(Bytes& o, const G1Affine& p) {
    if (p.is_inf()) { size_t s = o.size(); o.resize(s + 48, 0); o[s + 47] |= 1 << 6; return; }
    ser_fq_raw(o, p.x);
    if (Fq::cmp_canonical(p.y, Fq::neg(p.y)) > 0) o.back() |= 1 << 7;
}
static void ser_g2(Bytes& o, const G2Affine& p) {
    if (p.is_inf()) { size_t s = o.size(); o.resize(s + 96, 0); o[s + 95] |= 1 << 6; return; }
    ser_fq_raw(o, p.x.c0); ser_fq_raw(o, p.x.c1);
    if (Fq2::cmp_canonical(p.y, Fq2::neg(p.y)) > 0) o.back() |= 1 << 7;
}
// MatrixExtension { constraint: Vec<Vec<(F, usize)>>, num_constraints: usize }  r1cs_reader.rs:9-13
static void ser_matrix(Bytes& o, const Matrix& m, size_t num_constraints) {
    ser_u64(o, m.size());
    for (auto& row : m) {
        ser_u64(o, row.size());
        for (auto& e : row) { ser_fr(o, e.first); ser_u64(o, e.second); }
    }
    ser_u64(o, num_constraints);
}
// UPSTREAM linear_sumcheck IndexInfo { max_multiplicands, num_variables } (field order UNVERIFIED)
static void ser_index_info(Bytes& o, size_t max_multiplicands, size_t num_variables) {
    ser_u64(o, max_multiplicands); ser_u64(o, num_variables);
}

// ------------------------------------------------------------------ MLExtensionArray (UPSTREAM)
// fix the lowest variable: T'[b] = T[2b] (1 - r) + T[2b+1] r
static void mle_fold(std::vector<Fr>& t, const Fr& r) {
    size_t half = t.size() / 2;
    Fr omr = Fr::sub(Fr::R1, r);
    if (g_oracle_threads > 1 && half >= 512) {       // test-only threading (see g_oracle_threads): out of place
        std::vector<Fr> o(half);
#pragma omp parallel for num_threads(g_oracle_threads) schedule(static)
        for (size_t b = 0; b < half; b++) o[b] = Fr::add(Fr::mul(t[2 * b], omr), Fr::mul(t[2 * b + 1], r));
        t.swap(o);
        return;
    }
    for (size_t b = 0; b < half; b++)
        t[b] = Fr::add(Fr::mul(t[2 * b], omr), Fr::mul(t[2 * b + 1], r));
    t.resize(half);
}
static Fr mle_eval_at(std::vector<Fr> t, const std::vector<Fr>& point) {
    for (auto& r : point) mle_fold(t, r);
    return t[0];
}

// src/data_structures/eq.rs:5-20
static std::vector<std::vector<Fr>> eq_extension(const std::vector<Fr>& t) {
    size_t dim = t.size();
    std::vector<std::vector<Fr>> result(dim);
#pragma omp parallel for num_threads(g_oracle_threads) if (g_oracle_threads > 1 && dim >= 8) schedule(dynamic, 1)
    for (size_t i = 0; i < dim; i++) {
        std::vector<Fr> poly; poly.reserve(size_t(1) << dim);
        for (size_t x = 0; x < (size_t(1) << dim); x++) {
            Fr xi = ((x >> i) & 1) ? Fr::R1 : Fr::ZERO;
            Fr ti = t[i];
            Fr ti_xi = Fr::mul(ti, xi);
            poly.push_back(Fr::add(Fr::sub(Fr::sub(Fr::add(ti_xi, ti_xi), xi), ti), Fr::R1));
        }
        result[i] = std::move(poly);
    }
    return result;
}

// src/data_structures/r1cs_reader.rs:75-85
static std::vector<Fr> sum_over_y(const Matrix& m, const std::vector<Fr>& z) {
    std::vector<Fr> out(m.size());
    for (size_t x = 0; x < m.size(); x++) {
        Fr acc = Fr::ZERO;
        for (auto& e : m[x]) acc = Fr::add(acc, Fr::mul(e.first, z[e.second]));
        out[x] = acc;
    }
    return out;
}

// src/data_structures/r1cs_reader.rs:91-117: flatten to keys (y << s) + x, fix the low s variables
// one at a time (UPSTREAM SparseMLExtensionMap::eval_partial_at: entries with the same high part
// merge), scatter the survivors to a dense table.
static std::vector<Fr> eval_on_x(const Matrix& m, const std::vector<Fr>& r_x) {
    size_t s = r_x.size();
    std::unordered_map<uint64_t, Fr> cur;
    for (size_t x = 0; x < m.size(); x++)
        for (auto& e : m[x]) cur[((uint64_t)e.second << s) + x] = e.first;
    for (size_t i = 0; i < s; i++) {
        Fr r = r_x[i], omr = Fr::sub(Fr::R1, r);
        std::unordered_map<uint64_t, Fr> nxt;
        nxt.reserve(cur.size());
        for (auto& kv : cur) {
            Fr contrib = Fr::mul(kv.second, (kv.first & 1) ? r : omr);
            auto it = nxt.find(kv.first >> 1);
            if (it == nxt.end()) nxt.emplace(kv.first >> 1, contrib);
            else it->second = Fr::add(it->second, contrib);
        }
        cur.swap(nxt);
    }
    std::vector<Fr> ans(size_t(1) << s, Fr::ZERO);
    for (auto& kv : cur) ans[kv.first] = kv.second;
    return ans;
}

// ------------------------------------------------------------------ UPSTREAM AHPForMLSumcheck prover
struct MLSumcheckProver {
    std::vector<std::vector<std::vector<Fr>>> products;   // products[p][k] = table
    size_t nv, max_multiplicands, round;
    std::vector<Fr> randomness;
    void init(size_t nv_) {
        nv = nv_; round = 0; max_multiplicands = 0;
        for (auto& p : products) if (p.size() > max_multiplicands) max_multiplicands = p.size();
    }
    // prove_round(state, &Option<VerifierMsg>)
    std::vector<Fr> prove_round(const Fr* v_msg) {
        if (v_msg) {
            randomness.push_back(*v_msg);
            for (auto& p : products) for (auto& tab : p) mle_fold(tab, *v_msg);
        }
        round++;
        size_t half = size_t(1) << (nv - round);
        std::vector<Fr> sums(max_multiplicands + 1, Fr::ZERO);
        if (g_oracle_threads > 1 && half >= 256) {       // test-only threading: per-thread partial sums, added in order
            const int T = g_oracle_threads;
            std::vector<std::vector<Fr>> part(T, std::vector<Fr>(max_multiplicands + 1, Fr::ZERO));
#pragma omp parallel for num_threads(T) schedule(static)
            for (int k = 0; k < T; k++) round_range(half * k / T, half * (k + 1) / T, part[k]);
            for (int k = 0; k < T; k++) for (size_t t = 0; t <= max_multiplicands; t++) sums[t] = Fr::add(sums[t], part[k][t]);
            return sums;
        }
        round_range(0, half, sums);
        return sums;
    }
    void round_range(size_t b_lo, size_t b_hi, std::vector<Fr>& sums) const {
        for (size_t b = b_lo; b < b_hi; b++) {
            Fr t_as_field = Fr::ZERO;
            for (size_t t = 0; t <= max_multiplicands; t++) {
                Fr one_minus_t = Fr::sub(Fr::R1, t_as_field);
                for (auto& p : products) {
                    Fr product = Fr::R1;
                    for (auto& tab : p) {
                        Fr val = Fr::add(Fr::mul(tab[b << 1], one_minus_t), Fr::mul(tab[(b << 1) + 1], t_as_field));
                        product = Fr::mul(product, val);
                    }
                    sums[t] = Fr::add(sums[t], product);
                }
                t_as_field = Fr::add(t_as_field, Fr::R1);
            }
        }
    }
};

// ------------------------------------------------------------------ commitment (src/commitment/)
struct PublicParameter {            // data_structures.rs:10-17
    size_t nv;
    std::vector<std::vector<G1Affine>> powers_of_g;
    std::vector<std::vector<G2Affine>> powers_of_h;
    G1Affine g; G2Affine h;
    std::vector<Fr> trapdoor;       // kept for the identity tests (keygen returns it, setup.rs:104)
    std::vector<G1Affine> g_mask_random;   // VerifierParameter (data_structures.rs:20-25)
};

static void scalars_into_repr(const std::vector<Fr>& v, std::vector<uint64_t>& out) {
    out.resize(v.size() * 4);
    for (size_t i = 0; i < v.size(); i++) v[i].to_canonical(&out[4 * i]);
}

// setup.rs:27-105 with caller-supplied g, h, t (the reference draws them from its rng)
static void keygen(PublicParameter& pp, size_t nv, const G1Affine& g, const G2Affine& h, const std::vector<Fr>& t) {
    pp.nv = nv; pp.g = g; pp.h = h; pp.trapdoor = t;
    std::vector<Fr> pp_powers;
    for (size_t i = 0; i < nv; i++) {
        // eq(t[i..], x): variable k of x pairs with t[i + k]   (setup.rs:37-48 via remove_dummy_variable)
        std::vector<Fr> eq(1, Fr::R1);
        for (size_t k = nv; k-- > i;) {
            std::vector<Fr> nxt(eq.size() * 2);
            Fr tk = t[k], omt = Fr::sub(Fr::R1, tk);
            for (size_t b = 0; b < eq.size(); b++) { nxt[2 * b] = Fr::mul(eq[b], omt); nxt[2 * b + 1] = Fr::mul(eq[b], tk); }
            eq.swap(nxt);
        }
        pp_powers.insert(pp_powers.end(), eq.begin(), eq.end());
    }
    std::vector<uint64_t> sc; scalars_into_repr(pp_powers, sc);
    std::vector<G1Affine> pg; std::vector<G2Affine> ph;
    fixed_base_mul<Fq>(g, sc.data(), pp_powers.size(), pg);
    fixed_base_mul<Fq2>(h, sc.data(), pp_powers.size(), ph);
    size_t start = 0;
    pp.powers_of_g.clear(); pp.powers_of_h.clear();
    for (size_t i = 0; i < nv; i++) {
        size_t size = size_t(1) << (nv - i);
        pp.powers_of_g.emplace_back(pg.begin() + start, pg.begin() + start + size);
        pp.powers_of_h.emplace_back(ph.begin() + start, ph.begin() + start + size);
        start += size;
    }
    std::vector<uint64_t> tc; scalars_into_repr(t, tc);
    fixed_base_mul<Fq>(g, tc.data(), nv, pp.g_mask_random);
}

// commit.rs:17-29
static G1Affine pc_commit(const PublicParameter& pp, const std::vector<Fr>& poly) {
    std::vector<uint64_t> sc; scalars_into_repr(poly, sc);
    size_t n = std::min(poly.size(), pp.powers_of_g[0].size());
    return msm_pippenger<Fq>(pp.powers_of_g[0].data(), sc.data(), n).to_affine();
}

// open.rs:19-58
static Fr pc_open(const PublicParameter& pp, const std::vector<Fr>& poly, const std::vector<Fr>& point,
                  std::vector<G2Affine>& proofs, std::vector<std::vector<Fr>>* q_out = nullptr) {
    Fr eval_result = mle_eval_at(poly, point);
    size_t nv = point.size();
    std::vector<std::vector<Fr>> r(nv + 1), q(nv + 1);
    r[nv] = poly;
    proofs.clear();
    for (size_t i = 0; i < nv; i++) {
        size_t k = nv - i;
        Fr point_at_k = point[i];
        q[k].assign(size_t(1) << (k - 1), Fr::ZERO);
        r[k - 1].assign(size_t(1) << (k - 1), Fr::ZERO);
        Fr omp = Fr::sub(Fr::R1, point_at_k);
#pragma omp parallel for num_threads(g_oracle_threads) if (g_oracle_threads > 1 && k >= 10) schedule(static)
        for (size_t b = 0; b < (size_t(1) << (k - 1)); b++) {
            q[k][b] = Fr::sub(r[k][(b << 1) + 1], r[k][b << 1]);
            r[k - 1][b] = Fr::add(Fr::mul(r[k][b << 1], omp), Fr::mul(r[k][(b << 1) + 1], point_at_k));
        }
        std::vector<uint64_t> scalars((size_t(4)) << k);
#pragma omp parallel for num_threads(g_oracle_threads) if (g_oracle_threads > 1 && k >= 10) schedule(static)
        for (size_t x = 0; x < (size_t(1) << k); x++) q[k][x >> 1].to_canonical(&scalars[4 * x]);
        proofs.push_back(msm_pippenger<Fq2>(pp.powers_of_h[i].data(), scalars.data(), size_t(1) << k).to_affine());
        r[k].clear(); r[k].shrink_to_fit();
    }
    if (q_out) *q_out = q;
    return eval_result;
}

// ------------------------------------------------------------------ workload (constraints.rs + test_utils.rs)
struct R1CS {
    size_t log_n, n;
    Matrix a, b, c;
    std::vector<Fr> v, w;
};

// src/data_structures/constraints.rs:39-110 driven by src/test_utils.rs:51-102 (pad_to_square = true).
// Columns: instance variable i -> i (Variable::One = 0), witness j -> num_public + j (UPSTREAM to_matrices);
// LCs are kept sorted by variable with duplicates merged (UPSTREAM `lc + var`).
static void synth_r1cs(R1CS& cs, size_t num_public, size_t num_private, unsigned density, uint64_t seed) {
    SplitMix64 rng{seed};
    size_t n_inst = 1, n_wit = 0;          // Instance(0) = One
    std::vector<Fr> inst_val(1, Fr::R1), wit_val;
    std::vector<std::pair<Fr, size_t>> assignments;   // (value, column)
    auto new_input = [&](const Fr& v) { inst_val.push_back(v); return n_inst++; };
    std::vector<std::map<size_t, Fr>> ra, rb, rc;
    auto lc_add = [](std::map<size_t, Fr>& lc, size_t var) {
        auto it = lc.find(var);
        if (it == lc.end()) lc[var] = Fr::R1; else it->second = Fr::add(it->second, Fr::R1);
    };
    Fr a_val = fr_rand(rng); size_t a_var = new_input(a_val);
    assignments.push_back({a_val, a_var});
    Fr b_val = fr_rand(rng); size_t b_var = new_input(b_val);
    assignments.push_back({a_val, a_var});             // sic: constraints.rs:47 pushes (a_val, a_var) again
    for (size_t i = 0; i + 3 < num_public; i++) {
        Fr val = fr_rand(rng); size_t var = new_input(val);
        assignments.push_back({val, var});
    }
    // witness columns are offset by num_public once known; record as (1 << 62) + j until the end
    const size_t WIT = size_t(1) << 62;
    auto new_witness = [&](const Fr& v) { wit_val.push_back(v); return WIT + n_wit++; };
    size_t num_sparse = (num_private - 1) * (510 - density) / 510;
    for (size_t i = 0; i < num_sparse; i++) {
        size_t offset_var_index = 2 + (size_t)(rng.next_u64() % (uint64_t)(num_public - 1 - 2));   // gen_range(2, num_public - 1)
        Fr offset_val = assignments[offset_var_index].first; size_t offset_var = assignments[offset_var_index].second;
        std::map<size_t, Fr> la, lb, lc;
        Fr c_val; size_t c_var;
        if (i % 2 != 0) {
            c_val = Fr::mul(a_val, Fr::add(b_val, offset_val));
            c_var = new_witness(c_val);
            lc_add(la, a_var); lc_add(lb, b_var); lc_add(lb, offset_var); lc_add(lc, c_var);
        } else {
            c_val = Fr::add(Fr::add(a_val, b_val), offset_val);
            c_var = new_witness(c_val);
            lc_add(la, a_var); lc_add(la, b_var); lc_add(la, offset_var); lc_add(lb, 0); lc_add(lc, c_var);
        }
        ra.push_back(la); rb.push_back(lb); rc.push_back(lc);
        assignments.push_back({c_val, c_var});
        a_val = b_val; a_var = b_var; b_val = c_val; b_var = c_var;
    }
    for (size_t i = num_sparse; i < num_private; i++) {
        std::map<size_t, Fr> la, lb, lc;
        Fr c_val = Fr::ZERO;
        for (auto& as : assignments) { lc_add(la, as.second); lc_add(lb, as.second); c_val = Fr::add(c_val, as.first); }
        c_val = Fr::sqr(c_val);
        size_t c_var = new_witness(c_val);
        lc_add(lc, c_var);
        ra.push_back(la); rb.push_back(lb); rc.push_back(lc);
    }
    // make_matrices_square (test_utils.rs:81-102)
    size_t num_formatted = num_public + num_private, num_constraints = ra.size();
    if (num_formatted > num_constraints) {
        for (size_t i = num_constraints; i < num_formatted; i++) { ra.push_back({}); rb.push_back({}); rc.push_back({}); }
    } else {
        for (size_t i = num_formatted; i < num_constraints; i++) new_witness(Fr::R1);
    }
    auto finish = [&](std::vector<std::map<size_t, Fr>>& rows, Matrix& out) {
        out.clear(); out.resize(rows.size());
        for (size_t r = 0; r < rows.size(); r++)
            for (auto& kv : rows[r]) {
                if (kv.second.is_zero()) continue;
                size_t col = kv.first >= WIT ? (kv.first - WIT) + n_inst : kv.first;
                out[r].push_back({kv.second, col});
            }
    };
    // std::map orders WIT+j after every instance column, matching Variable's ordering
    finish(ra, cs.a); finish(rb, cs.b); finish(rc, cs.c);
    cs.v = inst_val; cs.w = wit_val;
    cs.n = cs.a.size();
    cs.log_n = 0; while ((size_t(1) << cs.log_n) < cs.n) cs.log_n++;
}

// ------------------------------------------------------------------ the NI prover (src/lib.rs:58-146)
struct Trace {
    std::map<std::string, Bytes> blobs;
    std::map<std::string, double> times;
    void put_fr(const std::string& k, const std::vector<Fr>& v) {
        Bytes& b = blobs[k]; b.resize(v.size() * 32);
        if (!v.empty()) memcpy(b.data(), v.data(), b.size());
    }
    void append_fr(const std::string& k, const std::vector<Fr>& v) {
        Bytes& b = blobs[k]; size_t s = b.size(); b.resize(s + v.size() * 32);
        if (!v.empty()) memcpy(b.data() + s, v.data(), v.size() * 32);
    }
};

static int prove(const R1CS& cs, const PublicParameter& pp, const std::vector<Fr>& v, const std::vector<Fr>& w,
                 Bytes& proof, Trace& tr) {
    size_t log_n = cs.log_n, n = cs.n;
    // prover_init (prover.rs:109-121)
    if (v.empty() || (v.size() & (v.size() - 1))) return 1;
    if (v.size() + w.size() != n) return 1;
    double t0 = now_s(), t_all = t0;
    FsRng fs; fs.setup();
    { Bytes b; ser_matrix(b, cs.a, n); fs.feed(b.data(), b.size()); }
    { Bytes b; ser_matrix(b, cs.b, n); fs.feed(b.data(), b.size()); }
    { Bytes b; ser_matrix(b, cs.c, n); fs.feed(b.data(), b.size()); }
    { Bytes b; ser_fr_vec(b, v); fs.feed(b.data(), b.size()); }
    tr.times["transcript_init"] = now_s() - t0;
    size_t log_v = 0; while ((size_t(1) << log_v) < v.size()) log_v++;

    // Prove 1: commitment (prover.rs:123-141)
    t0 = now_s();
    std::vector<Fr> z(v); z.insert(z.end(), w.begin(), w.end());
    G1Affine com = pc_commit(pp, z);
    tr.times["prove1_commit"] = now_s() - t0;
    Bytes pm1; ser_u64(pm1, log_n); ser_g1(pm1, com);
    fs.feed(pm1.data(), pm1.size());
    std::vector<Fr> r_v; for (size_t i = 0; i < log_v; i++) r_v.push_back(fr_rand(fs));   // verifier.rs:172-178

    // Prove 2: open at (r_v, 0, ..., 0) (prover.rs:143-160)
    t0 = now_s();
    std::vector<Fr> r_v0(r_v); r_v0.resize(log_n, Fr::ZERO);
    std::vector<G2Affine> proofs1;
    Fr z_rv_0 = pc_open(pp, z, r_v0, proofs1);
    tr.times["prove2_open"] = now_s() - t0;
    Bytes pm2; ser_fr(pm2, z_rv_0); ser_g2(pm2, pp.h); ser_u64(pm2, proofs1.size());
    for (auto& p : proofs1) ser_g2(pm2, p);
    fs.feed(pm2.data(), pm2.size());
    std::vector<Fr> tor; for (size_t i = 0; i < log_n; i++) tor.push_back(fr_rand(fs));   // verifier.rs:211-217

    // Prove 3 (prover.rs:163-196)
    t0 = now_s();
    std::vector<std::vector<Fr>> eq = eq_extension(tor);
    std::vector<Fr> az = sum_over_y(cs.a, z), bz = sum_over_y(cs.b, z), cz = sum_over_y(cs.c, z);
    MLSumcheckProver sc1;
    {
        std::vector<std::vector<Fr>> first; first.push_back(az); first.push_back(bz);
        for (auto& e : eq) first.push_back(e);
        std::vector<Fr> ncz(cz.size()); for (size_t i = 0; i < cz.size(); i++) ncz[i] = Fr::neg(cz[i]);
        std::vector<std::vector<Fr>> second; second.push_back(ncz);
        for (auto& e : eq) second.push_back(e);
        sc1.products.push_back(std::move(first)); sc1.products.push_back(std::move(second));
    }
    eq.clear();
    sc1.init(log_n);
    tr.times["prove3_setup"] = now_s() - t0;
    Bytes pm3; ser_index_info(pm3, sc1.max_multiplicands, log_n);
    fs.feed(pm3.data(), pm3.size());
    // Prove Sumcheck 1 (lib.rs:88-103)
    t0 = now_s();
    std::vector<Bytes> sc1_msgs;
    Fr vm; bool have_vm = false;
    for (size_t i = 0; i < log_n; i++) {
        std::vector<Fr> evals = sc1.prove_round(have_vm ? &vm : nullptr);
        tr.append_fr("sc1_evals", evals);
        Bytes pm; ser_fr_vec(pm, evals);
        fs.feed(pm.data(), pm.size()); sc1_msgs.push_back(pm);
        vm = fr_rand(fs); have_vm = true;
    }
    tr.times["sumcheck1"] = now_s() - t0;
    // Prove 4 (prover.rs:210-228)
    t0 = now_s();
    std::vector<Fr> r_x = sc1.randomness; r_x.push_back(vm);
    Fr va = mle_eval_at(az, r_x), vb = mle_eval_at(bz, r_x), vc = mle_eval_at(cz, r_x);
    sc1.products.clear();
    tr.times["prove4"] = now_s() - t0;
    Bytes pm4; ser_fr(pm4, va); ser_fr(pm4, vb); ser_fr(pm4, vc);
    fs.feed(pm4.data(), pm4.size());
    Fr r_a = fr_rand(fs), r_b = fr_rand(fs), r_c = fr_rand(fs);     // verifier.rs:354-360
    // Prove 5 (prover.rs:230-255)
    t0 = now_s();
    MLSumcheckProver sc2;
    std::vector<Fr> m_comb(n, Fr::ZERO);
    {
        const Matrix* ms[3] = {&cs.a, &cs.b, &cs.c}; Fr rk[3] = {r_a, r_b, r_c};
        std::vector<Fr> ev[3];
#pragma omp parallel for num_threads(g_oracle_threads < 3 ? g_oracle_threads : 3) if (g_oracle_threads > 1) schedule(static, 1)
        for (int k = 0; k < 3; k++) ev[k] = eval_on_x(*ms[k], r_x);
        for (int k = 0; k < 3; k++) {
            std::vector<Fr> t = std::move(ev[k]);
            for (auto& x : t) x = Fr::mul(x, rk[k]);                 // .multiply(r_k)
            for (size_t i = 0; i < n; i++) m_comb[i] = Fr::add(m_comb[i], t[i]);
            std::vector<std::vector<Fr>> prod; prod.push_back(std::move(t)); prod.push_back(z);
            sc2.products.push_back(std::move(prod));
        }
    }
    sc2.init(log_n);
    tr.times["prove5_eval_on_x"] = now_s() - t0;
    Bytes pm5; ser_index_info(pm5, sc2.max_multiplicands, log_n);
    fs.feed(pm5.data(), pm5.size());
    // Prove Sumcheck 2 (lib.rs:116-131)
    t0 = now_s();
    std::vector<Bytes> sc2_msgs; have_vm = false;
    for (size_t i = 0; i < log_n; i++) {
        std::vector<Fr> evals = sc2.prove_round(have_vm ? &vm : nullptr);
        tr.append_fr("sc2_evals", evals);
        Bytes pm; ser_fr_vec(pm, evals);
        fs.feed(pm.data(), pm.size()); sc2_msgs.push_back(pm);
        vm = fr_rand(fs); have_vm = true;
    }
    tr.times["sumcheck2"] = now_s() - t0;
    // Prove 6 (prover.rs:268-281)
    t0 = now_s();
    std::vector<Fr> r_y = sc2.randomness; r_y.push_back(vm);
    sc2.products.clear();
    std::vector<G2Affine> proofs2;
    Fr z_ry = pc_open(pp, z, r_y, proofs2);
    tr.times["prove6_open"] = now_s() - t0;
    Bytes pm6; ser_fr(pm6, z_ry); ser_g2(pm6, pp.h); ser_u64(pm6, proofs2.size());
    for (auto& p : proofs2) ser_g2(pm6, p);
    // Proof { .. } field order: src/data_structures/proof.rs:11-20
    proof.clear();
    proof.insert(proof.end(), pm1.begin(), pm1.end());
    proof.insert(proof.end(), pm2.begin(), pm2.end());
    proof.insert(proof.end(), pm3.begin(), pm3.end());
    ser_u64(proof, sc1_msgs.size()); for (auto& m : sc1_msgs) proof.insert(proof.end(), m.begin(), m.end());
    proof.insert(proof.end(), pm4.begin(), pm4.end());
    proof.insert(proof.end(), pm5.begin(), pm5.end());
    ser_u64(proof, sc2_msgs.size()); for (auto& m : sc2_msgs) proof.insert(proof.end(), m.begin(), m.end());
    proof.insert(proof.end(), pm6.begin(), pm6.end());
    tr.times["total"] = now_s() - t_all;

    tr.put_fr("az", az); tr.put_fr("bz", bz); tr.put_fr("cz", cz);
    tr.put_fr("r_v", r_v); tr.put_fr("tor", tor); tr.put_fr("r_x", r_x); tr.put_fr("r_y", r_y);
    tr.put_fr("vabc", {va, vb, vc}); tr.put_fr("r_abc", {r_a, r_b, r_c});
    tr.put_fr("z_rv_0", {z_rv_0}); tr.put_fr("z_ry", {z_ry}); tr.put_fr("m_comb", m_comb);
    { Bytes& b = tr.blobs["commitment"]; b.resize(96); memcpy(b.data(), &com, 96); }
    { Bytes& b = tr.blobs["open1_proofs"]; b.resize(192 * proofs1.size()); memcpy(b.data(), proofs1.data(), b.size()); }
    { Bytes& b = tr.blobs["open2_proofs"]; b.resize(192 * proofs2.size()); memcpy(b.data(), proofs2.data(), b.size()); }
    tr.blobs["proof"] = proof;
    return 0;
}


// ------------------------------------------------------------------ deserialization (UPSTREAM CanonicalDeserialize)
struct Reader {
    const uint8_t* p; size_t left; bool ok;
    Reader(const uint8_t* d, size_t n) : p(d), left(n), ok(true) {}
    bool take(void* out, size_t n) { if (left < n) { ok = false; return false; } memcpy(out, p, n); p += n; left -= n; return true; }
    uint64_t u64() { uint8_t b[8] = {0}; take(b, 8); uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | b[i]; return v; }
    Fr fr() {
        uint64_t c[4] = {0, 0, 0, 0}; uint8_t b[32] = {0}; take(b, 32);
        for (int i = 0; i < 4; i++) for (int j = 7; j >= 0; j--) c[i] = (c[i] << 8) | b[8 * i + j];
        if (Fr::geq_mod(c)) ok = false;
        return Fr::from_canonical(c);
    }
};
static bool fq_from_bytes(const uint8_t* b, Fq& out) {
    uint64_t c[6];
    for (int i = 0; i < 6; i++) { c[i] = 0; for (int j = 7; j >= 0; j--) c[i] = (c[i] << 8) | b[8 * i + j]; }
    if (Fq::geq_mod(c)) return false;
    out = Fq::from_canonical(c); return true;
}
static bool fq_sqrt(const Fq& a, Fq& out) {             // p = 3 mod 4: a^((p+1)/4)
    static const uint64_t E[6] = {0xee7fbfffffffeaabULL, 0x07aaffffac54ffffULL, 0xd9cc34a83dac3d89ULL, 0xd91dd2e13ce144afULL, 0x92c6e9ed90d2eb35ULL, 0x0680447a8e5ff9a6ULL};
    Fq s = Fq::pow(a, E, 6);
    if (Fq::sqr(s) != a) return false;
    out = s; return true;
}
static bool fq2_sqrt(const Fq2& a, Fq2& out) {          // complex method
    if (a.is_zero()) { out = a; return true; }
    Fq n = Fq::add(Fq::sqr(a.c0), Fq::sqr(a.c1)), s;
    if (!fq_sqrt(n, s)) return false;
    Fq inv2 = Fq::inv(Fq::from_u64(2));
    for (int k = 0; k < 2; k++) {
        Fq sg = k == 0 ? s : Fq::neg(s);
        Fq d = Fq::mul(Fq::add(a.c0, sg), inv2), x0;
        if (!fq_sqrt(d, x0) || x0.is_zero()) continue;
        Fq x1 = Fq::mul(a.c1, Fq::inv(Fq::dbl(x0)));
        Fq2 c; c.c0 = x0; c.c1 = x1;
        if (Fq2::sqr(c) == a) { out = c; return true; }
    }
    return false;
}
static bool read_g1(Reader& r, G1Affine& out) {
    uint8_t b[48]; if (!r.take(b, 48)) return false;
    bool inf = b[47] & 0x40, pos = b[47] & 0x80; b[47] &= 0x3f;
    if (inf) { out = G1Affine::inf(); return true; }
    Fq x, y;
    if (!fq_from_bytes(b, x)) return r.ok = false;
    if (!fq_sqrt(Fq::add(Fq::mul(Fq::sqr(x), x), Fq::from_u64(4)), y)) return r.ok = false;
    bool y_larger = Fq::cmp_canonical(y, Fq::neg(y)) > 0;
    out.x = x; out.y = (y_larger == pos) ? y : Fq::neg(y);
    return true;
}
static bool read_g2(Reader& r, G2Affine& out) {
    uint8_t b[96]; if (!r.take(b, 96)) return false;
    bool inf = b[95] & 0x40, pos = b[95] & 0x80; b[95] &= 0x3f;
    if (inf) { out = G2Affine::inf(); return true; }
    Fq2 x, y;
    if (!fq_from_bytes(b, x.c0) || !fq_from_bytes(b + 48, x.c1)) return r.ok = false;
    Fq2 bb; bb.c0 = Fq::from_u64(4); bb.c1 = Fq::from_u64(4);
    if (!fq2_sqrt(Fq2::add(Fq2::mul(Fq2::sqr(x), x), bb), y)) return r.ok = false;
    bool y_larger = Fq2::cmp_canonical(y, Fq2::neg(y)) > 0;
    out.x = x; out.y = (y_larger == pos) ? y : Fq2::neg(y);
    return true;
}

// ------------------------------------------------------------------ verifier
// src/commitment/verify.rs:12-45:  e(C - g^eval, h) == prod_i e(g^{t_i} - g^{p_i}, pi_i)
static bool pc_verify(const PublicParameter& vp, const G1Affine& commitment, const std::vector<Fr>& point, const Fr& eval,
                      const std::vector<G2Affine>& proofs) {
    if (proofs.size() != vp.nv || point.size() != vp.nv) return false;
    uint64_t c[4];
    eval.to_canonical(c);
    G1Jac left = G1Jac::add(G1Jac::from_affine(commitment), G1Jac::neg(G1Jac::mul(G1Jac::from_affine(vp.g), c, 4)));
    Fq12 lhs = miller_loop(left.to_affine(), vp.h);
    Fq12 rhs = Fq12::one();
    for (size_t i = 0; i < vp.nv; i++) {
        point[i].to_canonical(c);
        G1Jac l = G1Jac::add(G1Jac::from_affine(vp.g_mask_random[i]), G1Jac::neg(G1Jac::mul(G1Jac::from_affine(vp.g), c, 4)));
        rhs = Fq12::mul(rhs, miller_loop(l.to_affine(), proofs[i]));
    }
    return final_exponentiation(lhs) == final_exponentiation(rhs);
}
// UPSTREAM interpolate_uni_poly: value at r of the polynomial with the given evaluations at 0..d
static Fr interpolate(const std::vector<Fr>& ev, const Fr& r) {
    size_t d = ev.size() - 1;
    std::vector<Fr> diff(d + 1);
    Fr prod = Fr::R1;
    for (size_t k = 0; k <= d; k++) {
        diff[k] = Fr::sub(r, Fr::from_u64(k));
        if (diff[k].is_zero()) return ev[k];
        prod = Fr::mul(prod, diff[k]);
    }
    Fr acc = Fr::ZERO;
    for (size_t i = 0; i <= d; i++) {
        Fr w = Fr::R1;                                  // prod_{k != i} (i - k)
        for (size_t k = 0; k <= d; k++) if (k != i) w = Fr::mul(w, Fr::sub(Fr::from_u64(i), Fr::from_u64(k)));
        acc = Fr::add(acc, Fr::mul(ev[i], Fr::mul(prod, Fr::inv(Fr::mul(w, diff[i])))));
    }
    return acc;
}
// UPSTREAM check_and_generate_subclaim: returns false when some round violates P_j(0) + P_j(1) == expected_j
static bool sumcheck_subclaim(const std::vector<std::vector<Fr>>& msgs, const std::vector<Fr>& chal, const Fr& asserted, Fr& expected_out) {
    Fr expected = asserted;
    for (size_t j = 0; j < msgs.size(); j++) {
        if (msgs[j].size() < 2) return false;
        if (Fr::add(msgs[j][0], msgs[j][1]) != expected) return false;
        expected = interpolate(msgs[j], chal[j]);
    }
    expected_out = expected; return true;
}

// MLArgumentForR1CS::verify (src/lib.rs:147-212) + verify_sixth_round (src/ahp/verifier.rs:443-512).
// Returns 1 = accept; 0 = malformed proof; negative = the failed check.
static int verify(const R1CS& cs, const PublicParameter& vp, const std::vector<Fr>& v, const uint8_t* proof, size_t len) {
    const size_t log_n = cs.log_n, n = cs.n;
    if (v.empty() || (v.size() & (v.size() - 1)) || v.size() > n) return 0;       // verifier.rs:144-146
    size_t log_v = 0; while ((size_t(1) << log_v) < v.size()) log_v++;
    Reader rd(proof, len);
    FsRng fs; fs.setup();
    { Bytes b; ser_matrix(b, cs.a, n); fs.feed(b.data(), b.size()); }
    { Bytes b; ser_matrix(b, cs.b, n); fs.feed(b.data(), b.size()); }
    { Bytes b; ser_matrix(b, cs.c, n); fs.feed(b.data(), b.size()); }
    { Bytes b; ser_fr_vec(b, v); fs.feed(b.data(), b.size()); }
    auto feed_span = [&](const uint8_t* from) { fs.feed(from, (size_t)(rd.p - from)); };
    // pm1
    const uint8_t* mark = rd.p;
    uint64_t com_nv = rd.u64(); G1Affine com; if (!read_g1(rd, com)) return 0;
    (void)com_nv;
    feed_span(mark);
    std::vector<Fr> r_v; for (size_t i = 0; i < log_v; i++) r_v.push_back(fr_rand(fs));
    // pm2
    mark = rd.p;
    Fr z_rv_0 = rd.fr(); G2Affine h1; if (!read_g2(rd, h1)) return 0;
    uint64_t np1 = rd.u64(); if (!rd.ok || np1 != log_n) return 0;
    std::vector<G2Affine> proofs1(np1); for (auto& q : proofs1) if (!read_g2(rd, q)) return 0;
    feed_span(mark);
    std::vector<Fr> tor; for (size_t i = 0; i < log_n; i++) tor.push_back(fr_rand(fs));
    // pm3
    mark = rd.p;
    uint64_t mm1 = rd.u64(), nv1 = rd.u64();
    if (!rd.ok || nv1 != log_n) return 0;                                        // verifier.rs:249-252
    feed_span(mark);
    uint64_t cnt = rd.u64(); if (!rd.ok || cnt != log_n) return 0;
    std::vector<std::vector<Fr>> sc1(log_n); std::vector<Fr> r_x;
    for (size_t j = 0; j < log_n; j++) {
        mark = rd.p;
        uint64_t k = rd.u64(); if (!rd.ok || k != mm1 + 1) return 0;
        for (uint64_t t = 0; t < k; t++) sc1[j].push_back(rd.fr());
        if (!rd.ok) return 0;
        feed_span(mark);
        r_x.push_back(fr_rand(fs));
    }
    // pm4
    mark = rd.p;
    Fr va = rd.fr(), vb = rd.fr(), vc = rd.fr(); if (!rd.ok) return 0;
    feed_span(mark);
    Fr r_a = fr_rand(fs), r_b = fr_rand(fs), r_c = fr_rand(fs);
    // pm5
    mark = rd.p;
    uint64_t mm2 = rd.u64(), nv2 = rd.u64();
    if (!rd.ok || nv2 != log_n) return 0;                                        // verifier.rs:399-402
    feed_span(mark);
    cnt = rd.u64(); if (!rd.ok || cnt != log_n) return 0;
    std::vector<std::vector<Fr>> sc2(log_n); std::vector<Fr> r_y;
    for (size_t j = 0; j < log_n; j++) {
        mark = rd.p;
        uint64_t k = rd.u64(); if (!rd.ok || k != mm2 + 1) return 0;
        for (uint64_t t = 0; t < k; t++) sc2[j].push_back(rd.fr());
        if (!rd.ok) return 0;
        feed_span(mark);
        r_y.push_back(fr_rand(fs));
    }
    // pm6
    Fr z_ry = rd.fr(); G2Affine h2; if (!read_g2(rd, h2)) return 0;
    uint64_t np2 = rd.u64(); if (!rd.ok || np2 != log_n) return 0;
    std::vector<G2Affine> proofs2(np2); for (auto& q : proofs2) if (!read_g2(rd, q)) return 0;
    if (!rd.ok || rd.left != 0) return 0;

    // verify_sixth_round (verifier.rs:443-512)
    std::vector<Fr> r_v_0(r_v); r_v_0.resize(log_n, Fr::ZERO);
    if (!pc_verify(vp, com, r_v_0, z_rv_0, proofs1)) return -1;                  // "public witness failed in commitment check"
    if (mle_eval_at(v, r_v) != z_rv_0) return -2;                                // "public witness is inconsistent with proof"
    Fr expected1;
    if (!sumcheck_subclaim(sc1, r_x, Fr::ZERO, expected1)) return -3;
    {
        Fr eq_rx = Fr::R1;                                                       // prod_i eq_i(r_x) (verifier.rs:471-474)
        for (size_t i = 0; i < log_n; i++) {
            Fr t = tor[i], x = r_x[i];
            Fr tx = Fr::mul(t, x);
            eq_rx = Fr::mul(eq_rx, Fr::add(Fr::sub(Fr::sub(Fr::add(tx, tx), x), t), Fr::R1));
        }
        if (Fr::mul(Fr::sub(Fr::mul(va, vb), vc), eq_rx) != expected1) return -4; // "first sumcheck has wrong subclaim"
    }
    Fr claimed2 = Fr::add(Fr::add(Fr::mul(r_a, va), Fr::mul(r_b, vb)), Fr::mul(r_c, vc));
    Fr expected2;
    if (!sumcheck_subclaim(sc2, r_y, claimed2, expected2)) return -5;
    Fr a_rxy = mle_eval_at(eval_on_x(cs.a, r_x), r_y), b_rxy = mle_eval_at(eval_on_x(cs.b, r_x), r_y), c_rxy = mle_eval_at(eval_on_x(cs.c, r_x), r_y);
    Fr actual = Fr::mul(Fr::add(Fr::add(Fr::mul(r_a, a_rxy), Fr::mul(r_b, b_rxy)), Fr::mul(r_c, c_rxy)), z_ry);
    if (expected2 != actual) return -6;                                          // "Cannot verify matrix A, B, C"
    if (!pc_verify(vp, com, r_y, z_ry, proofs2)) return -7;                      // "Cannot verify z_ry"
    return 1;
}

// ------------------------------------------------------------------ generators
static bool hex_to_limbs(const char* hex, uint64_t* out, int n) {
    for (int i = 0; i < n; i++) out[i] = 0;
    size_t len = strlen(hex);
    for (size_t i = 0; i < len; i++) {
        char ch = hex[len - 1 - i]; int d;
        if (ch >= '0' && ch <= '9') d = ch - '0'; else if (ch >= 'a' && ch <= 'f') d = ch - 'a' + 10; else return false;
        if ((int)(i / 16) >= n) return false;
        out[i / 16] |= (uint64_t)d << (4 * (i % 16));
    }
    return true;
}
static Fq fq_from_hex(const char* hex) { uint64_t c[6]; hex_to_limbs(hex, c, 6); return Fq::from_canonical(c); }
// r-torsion generators (oracle/pymodel.py derive_generators(); G1 is the standard BLS12-381 generator)
static G1Affine g1_generator() {
    G1Affine g;
    g.x = fq_from_hex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
    g.y = fq_from_hex("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1");
    return g;
}
static G2Affine g2_generator() {
    G2Affine h;
    h.x.c0 = fq_from_hex("04d1cc4ad56b68cdb595adb46cad2cc82e3d0da9a75ef283b6bbd91df14533e1a45128ec26f8ab25072da969d7628b70");
    h.x.c1 = fq_from_hex("13a471d5149813b306fe76921cff7bb8d5c03fdc24a613f3e7a7fb8deb8097699751485a0bd2ad391718aaa4419ce75b");
    h.y.c0 = fq_from_hex("0a3d002cac5c50eb9e97e8b62ca30ffc5bf5aaacec121cdb63e19a5e358c4804439edb98366c02fd2840c7b9004f8b99");
    h.y.c1 = fq_from_hex("1834907430540701fa8aa597f79e63960ec77037a7d9a06606c4c58bd8019969edabb81b77fae18489a80d47bab79d25");
    return h;
}

// ================================================================== C interface (ctypes)
// Fr: 32-byte little-endian Montgomery image (== arkworks Fp256 memory).  G1 affine: x,y Fq Montgomery
// (96 B); G2 affine: x.c0,x.c1,y.c0,y.c1 (192 B); all-zero = infinity.
extern "C" {

int or_init() { ff_init_all(); return 0; }
// test-only: worker threads of the heavy loops (see g_oracle_threads in ec.hpp); returns the previous value
int or_set_threads(int k) { int old = g_oracle_threads; g_oracle_threads = k < 1 ? 1 : k; return old; }

void or_fr_binop(int op, const Fr* a, const Fr* b, Fr* out, size_t n) {
    ff_init_all();
    for (size_t i = 0; i < n; i++) out[i] = op == 0 ? Fr::add(a[i], b[i]) : op == 1 ? Fr::sub(a[i], b[i]) : Fr::mul(a[i], b[i]);
}
void or_fq_binop(int op, const Fq* a, const Fq* b, Fq* out, size_t n) {
    ff_init_all();
    for (size_t i = 0; i < n; i++) out[i] = op == 0 ? Fq::add(a[i], b[i]) : op == 1 ? Fq::sub(a[i], b[i]) : Fq::mul(a[i], b[i]);
}
void or_fr_from_canonical(const uint64_t* c, Fr* out, size_t n) { ff_init_all(); for (size_t i = 0; i < n; i++) out[i] = Fr::from_canonical(c + 4 * i); }
void or_fr_to_canonical(const Fr* a, uint64_t* out, size_t n) { ff_init_all(); for (size_t i = 0; i < n; i++) a[i].to_canonical(out + 4 * i); }
void or_fq_from_canonical(const uint64_t* c, Fq* out, size_t n) { ff_init_all(); for (size_t i = 0; i < n; i++) out[i] = Fq::from_canonical(c + 6 * i); }
void or_fq_to_canonical(const Fq* a, uint64_t* out, size_t n) { ff_init_all(); for (size_t i = 0; i < n; i++) a[i].to_canonical(out + 6 * i); }
void or_fr_rand(uint64_t seed, Fr* out, size_t n) { ff_init_all(); SplitMix64 rng{seed}; for (size_t i = 0; i < n; i++) out[i] = fr_rand(rng); }

void or_generators(G1Affine* g, G2Affine* h) { ff_init_all(); *g = g1_generator(); *h = g2_generator(); }

// ---- group helpers for tests
void or_g1_mul(const G1Affine* p, const Fr* k, G1Affine* out) {
    ff_init_all(); uint64_t c[4]; k->to_canonical(c);
    *out = G1Jac::mul(G1Jac::from_affine(*p), c, 4).to_affine();
}
void or_g2_mul(const G2Affine* p, const Fr* k, G2Affine* out) {
    ff_init_all(); uint64_t c[4]; k->to_canonical(c);
    *out = G2Jac::mul(G2Jac::from_affine(*p), c, 4).to_affine();
}
void or_g1_add(const G1Affine* a, const G1Affine* b, G1Affine* out) { ff_init_all(); *out = G1Jac::add_mixed(G1Jac::from_affine(*a), *b).to_affine(); }
void or_g2_add(const G2Affine* a, const G2Affine* b, G2Affine* out) { ff_init_all(); *out = G2Jac::add_mixed(G2Jac::from_affine(*a), *b).to_affine(); }
int or_g1_on_curve(const G1Affine* p) {
    ff_init_all(); if (p->is_inf()) return 1;
    Fq b = Fq::from_u64(4);
    return Fq::sqr(p->y) == Fq::add(Fq::mul(Fq::sqr(p->x), p->x), b);
}
int or_g2_on_curve(const G2Affine* p) {
    ff_init_all(); if (p->is_inf()) return 1;
    Fq2 b; b.c0 = Fq::from_u64(4); b.c1 = Fq::from_u64(4);
    return Fq2::sqr(p->y) == Fq2::add(Fq2::mul(Fq2::sqr(p->x), p->x), b);
}
void or_ser_g1(const G1Affine* p, uint8_t* out48) { ff_init_all(); Bytes b; ser_g1(b, *p); memcpy(out48, b.data(), 48); }
void or_ser_g2(const G2Affine* p, uint8_t* out96) { ff_init_all(); Bytes b; ser_g2(b, *p); memcpy(out96, b.data(), 96); }
void or_msm_g1(const G1Affine* bases, const Fr* scalars, size_t n, G1Affine* out) {
    ff_init_all(); std::vector<Fr> s(scalars, scalars + n); std::vector<uint64_t> sc; scalars_into_repr(s, sc);
    *out = msm_pippenger<Fq>(bases, sc.data(), n).to_affine();
}
void or_msm_g2(const G2Affine* bases, const Fr* scalars, size_t n, G2Affine* out) {
    ff_init_all(); std::vector<Fr> s(scalars, scalars + n); std::vector<uint64_t> sc; scalars_into_repr(s, sc);
    *out = msm_pippenger<Fq2>(bases, sc.data(), n).to_affine();
}

// ---- transcript
void* or_fs_new() { FsRng* f = new FsRng; f->setup(); return f; }
void or_fs_free(void* f) { delete (FsRng*)f; }
void or_fs_feed(void* f, const uint8_t* d, size_t n) { ((FsRng*)f)->feed(d, n); }
void or_fs_fill(void* f, uint8_t* d, size_t n) { ((FsRng*)f)->fill_bytes(d, n); }
void or_fs_fr_rand(void* f, Fr* out, size_t n) { ff_init_all(); for (size_t i = 0; i < n; i++) out[i] = fr_rand(*(FsRng*)f); }

// ---- R1CS
void* or_r1cs_synth(size_t num_public, size_t num_private, unsigned density, uint64_t seed) {
    ff_init_all(); R1CS* cs = new R1CS; synth_r1cs(*cs, num_public, num_private, density, seed); return cs;
}
void* or_r1cs_from_csr(size_t log_n, const uint64_t* const* rowptr, const uint32_t* const* col, const Fr* const* val) {
    ff_init_all(); R1CS* cs = new R1CS; cs->log_n = log_n; cs->n = size_t(1) << log_n;
    Matrix* ms[3] = {&cs->a, &cs->b, &cs->c};
    for (int k = 0; k < 3; k++) {
        ms[k]->resize(cs->n);
        for (size_t r = 0; r < cs->n; r++)
            for (uint64_t e = rowptr[k][r]; e < rowptr[k][r + 1]; e++) (*ms[k])[r].push_back({val[k][e], (size_t)col[k][e]});
    }
    return cs;
}
void or_r1cs_free(void* h) { delete (R1CS*)h; }
size_t or_r1cs_n(void* h) { return ((R1CS*)h)->n; }
size_t or_r1cs_num_public(void* h) { return ((R1CS*)h)->v.size(); }
size_t or_r1cs_nnz(void* h, int which) {
    R1CS* cs = (R1CS*)h; const Matrix& m = which == 0 ? cs->a : which == 1 ? cs->b : cs->c;
    size_t t = 0; for (auto& r : m) t += r.size(); return t;
}
void or_r1cs_export(void* h, int which, uint64_t* rowptr, uint32_t* col, Fr* val) {
    R1CS* cs = (R1CS*)h; const Matrix& m = which == 0 ? cs->a : which == 1 ? cs->b : cs->c;
    size_t e = 0;
    for (size_t r = 0; r < m.size(); r++) {
        rowptr[r] = e;
        for (auto& x : m[r]) { col[e] = (uint32_t)x.second; val[e] = x.first; e++; }
    }
    rowptr[m.size()] = e;
}
void or_r1cs_vw(void* h, Fr* v, Fr* w) {
    R1CS* cs = (R1CS*)h;
    if (v) memcpy(v, cs->v.data(), cs->v.size() * 32);
    if (w) memcpy(w, cs->w.data(), cs->w.size() * 32);
}
// 1 iff (Az) o (Bz) == Cz
int or_r1cs_is_satisfied(void* h, const Fr* z) {
    R1CS* cs = (R1CS*)h; std::vector<Fr> zz(z, z + cs->n);
    auto az = sum_over_y(cs->a, zz), bz = sum_over_y(cs->b, zz), cz = sum_over_y(cs->c, zz);
    for (size_t i = 0; i < cs->n; i++) if (Fr::mul(az[i], bz[i]) != cz[i]) return 0;
    return 1;
}
void or_sum_over_y(void* h, int which, const Fr* z, Fr* out) {
    R1CS* cs = (R1CS*)h; const Matrix& m = which == 0 ? cs->a : which == 1 ? cs->b : cs->c;
    std::vector<Fr> zz(z, z + cs->n); auto r = sum_over_y(m, zz); memcpy(out, r.data(), r.size() * 32);
}
void or_eval_on_x(void* h, int which, const Fr* r_x, Fr* out) {
    R1CS* cs = (R1CS*)h; const Matrix& m = which == 0 ? cs->a : which == 1 ? cs->b : cs->c;
    std::vector<Fr> rx(r_x, r_x + cs->log_n); auto r = eval_on_x(m, rx); memcpy(out, r.data(), r.size() * 32);
}
void or_eq_extension(const Fr* t, size_t dim, Fr* out /* dim * 2^dim */) {
    ff_init_all(); std::vector<Fr> tv(t, t + dim); auto e = eq_extension(tv);
    for (size_t i = 0; i < dim; i++) memcpy(out + (i << dim), e[i].data(), 32 << dim);
}
void or_mle_eval(const Fr* table, size_t nv, const Fr* point, Fr* out) {
    ff_init_all(); std::vector<Fr> t(table, table + (size_t(1) << nv)), p(point, point + nv); *out = mle_eval_at(t, p);
}
// one literal sumcheck proving run over caller tables: products given as table lists
// tables: ntab tables of 2^nv; prod_sizes[np]; prod_idx flattened.  challenges[nv] are consumed in order
// (challenge i is applied before round i+1).  out: nv * (max_mult + 1) evaluations.
void or_sumcheck_prove(const Fr* tables, size_t ntab, size_t nv, const uint32_t* prod_sizes, size_t np,
                       const uint32_t* prod_idx, const Fr* challenges, Fr* out) {
    ff_init_all(); MLSumcheckProver sc; size_t n = size_t(1) << nv; size_t k = 0;
    for (size_t p = 0; p < np; p++) {
        std::vector<std::vector<Fr>> prod;
        for (uint32_t j = 0; j < prod_sizes[p]; j++, k++) {
            const Fr* t = tables + (size_t)prod_idx[k] * n; prod.emplace_back(t, t + n);
        }
        sc.products.push_back(std::move(prod));
    }
    (void)ntab;
    sc.init(nv);
    size_t d = sc.max_multiplicands + 1;
    for (size_t i = 0; i < nv; i++) {
        auto ev = sc.prove_round(i ? &challenges[i - 1] : nullptr);
        memcpy(out + i * d, ev.data(), d * 32);
    }
}

// ---- public parameters
void* or_keygen(size_t nv, uint64_t seed) {
    ff_init_all(); SplitMix64 rng{seed};
    std::vector<Fr> t; for (size_t i = 0; i < nv; i++) t.push_back(fr_rand(rng));
    PublicParameter* pp = new PublicParameter; keygen(*pp, nv, g1_generator(), g2_generator(), t); return pp;
}
void* or_keygen_with(size_t nv, const G1Affine* g, const G2Affine* h, const Fr* t) {
    ff_init_all(); std::vector<Fr> tv(t, t + nv);
    PublicParameter* pp = new PublicParameter; keygen(*pp, nv, *g, *h, tv); return pp;
}
// build a PublicParameter from flat arrays: g1 = powers_of_g[0] (2^nv), g2 = all levels of powers_of_h
// concatenated (2^nv + 2^(nv-1) + ... + 2), h
void* or_pp_from_arrays(size_t nv, const G1Affine* g1_level0, const G2Affine* g2_all, const G2Affine* h) {
    ff_init_all(); PublicParameter* pp = new PublicParameter; pp->nv = nv; pp->h = *h; pp->g = G1Affine::inf();
    pp->powers_of_g.emplace_back(g1_level0, g1_level0 + (size_t(1) << nv));
    size_t start = 0;
    for (size_t i = 0; i < nv; i++) { size_t sz = size_t(1) << (nv - i); pp->powers_of_h.emplace_back(g2_all + start, g2_all + start + sz); start += sz; }
    return pp;
}
// verifier parameters only (VerifierParameter, data_structures.rs:20-25): enough for or_verify / or_pc_verify
void* or_vp_from_arrays(size_t nv, const G1Affine* g, const G2Affine* h, const G1Affine* g_mask) {
    ff_init_all(); PublicParameter* pp = new PublicParameter; pp->nv = nv; pp->g = *g; pp->h = *h;
    pp->g_mask_random.assign(g_mask, g_mask + nv);
    return pp;
}
void or_pp_free(void* h) { delete (PublicParameter*)h; }
void or_pp_export_g1(void* h, size_t level, G1Affine* out) { auto& v = ((PublicParameter*)h)->powers_of_g[level]; memcpy(out, v.data(), v.size() * sizeof(G1Affine)); }
void or_pp_export_g2(void* h, size_t level, G2Affine* out) { auto& v = ((PublicParameter*)h)->powers_of_h[level]; memcpy(out, v.data(), v.size() * sizeof(G2Affine)); }
void or_pp_gh(void* h, G1Affine* g, G2Affine* hh) { *g = ((PublicParameter*)h)->g; *hh = ((PublicParameter*)h)->h; }
void or_pp_trapdoor(void* h, Fr* out) { auto& t = ((PublicParameter*)h)->trapdoor; memcpy(out, t.data(), t.size() * 32); }
void or_pp_g_mask(void* h, G1Affine* out) { auto& t = ((PublicParameter*)h)->g_mask_random; memcpy(out, t.data(), t.size() * sizeof(G1Affine)); }

void or_commit(void* pph, const Fr* z, size_t n, G1Affine* out) {
    std::vector<Fr> zz(z, z + n); *out = pc_commit(*(PublicParameter*)pph, zz);
}
void or_open(void* pph, const Fr* z, const Fr* point, size_t nv, Fr* eval, G2Affine* proofs, Fr* q_flat /* nullable: q[nv], q[nv-1], .. q[1] */) {
    std::vector<Fr> zz(z, z + (size_t(1) << nv)), p(point, point + nv); std::vector<G2Affine> pr; std::vector<std::vector<Fr>> q;
    *eval = pc_open(*(PublicParameter*)pph, zz, p, pr, &q);
    memcpy(proofs, pr.data(), pr.size() * sizeof(G2Affine));
    if (q_flat) { size_t o = 0; for (size_t k = nv; k >= 1; k--) { memcpy(q_flat + o, q[k].data(), q[k].size() * 32); o += q[k].size(); } }
}


// ---- verifier
int or_verify(void* r1cs, void* vp, const Fr* v, size_t nv_len, const uint8_t* proof, size_t len) {
    std::vector<Fr> vv(v, v + nv_len);
    return verify(*(R1CS*)r1cs, *(PublicParameter*)vp, vv, proof, len);
}
int or_pc_verify(void* vp, const G1Affine* commitment, const Fr* point, const Fr* eval, const G2Affine* proofs) {
    PublicParameter* pp = (PublicParameter*)vp;
    std::vector<Fr> p(point, point + pp->nv); std::vector<G2Affine> pr(proofs, proofs + pp->nv);
    return pc_verify(*pp, *commitment, p, *eval, pr) ? 1 : 0;
}
// e(a P, b Q) == e(P, Q)^(ab) and e(P, Q) != 1: returns 1 when both hold
int or_pairing_check(const Fr* a, const Fr* b) {
    ff_init_all();
    G1Affine g = g1_generator(); G2Affine h = g2_generator();
    uint64_t ca[4], cb[4]; a->to_canonical(ca); b->to_canonical(cb);
    G1Affine ag = G1Jac::mul(G1Jac::from_affine(g), ca, 4).to_affine();
    G2Affine bh = G2Jac::mul(G2Jac::from_affine(h), cb, 4).to_affine();
    Fq12 e = pairing(g, h);
    if (e == Fq12::one()) return 0;
    uint64_t cab[4]; Fr::mul(*a, *b).to_canonical(cab);
    Fq12 acc = Fq12::one();
    for (int i = 255; i >= 0; i--) { acc = Fq12::sqr(acc); if ((cab[i / 64] >> (i % 64)) & 1) acc = Fq12::mul(acc, e); }
    return pairing(ag, bh) == acc ? 1 : 0;
}
// decompress round trip helpers
int or_deser_g1(const uint8_t* in48, G1Affine* out) { ff_init_all(); Reader r(in48, 48); return read_g1(r, *out) && r.ok; }
int or_deser_g2(const uint8_t* in96, G2Affine* out) { ff_init_all(); Reader r(in96, 96); return read_g2(r, *out) && r.ok; }

// ---- prove
void* or_prove(void* r1cs, void* pph, const Fr* v, size_t nv_len, const Fr* w, size_t nw_len, int* status) {
    Trace* tr = new Trace; Bytes proof;
    std::vector<Fr> vv(v, v + nv_len), ww(w, w + nw_len);
    *status = prove(*(R1CS*)r1cs, *(PublicParameter*)pph, vv, ww, proof, *tr);
    return tr;
}
void or_trace_free(void* t) { delete (Trace*)t; }
size_t or_trace_get(void* t, const char* name, const uint8_t** ptr) {
    Trace* tr = (Trace*)t; auto it = tr->blobs.find(name);
    if (it == tr->blobs.end()) { *ptr = nullptr; return 0; }
    *ptr = it->second.data(); return it->second.size();
}
double or_trace_time(void* t, const char* name) {
    Trace* tr = (Trace*)t; auto it = tr->times.find(name); return it == tr->times.end() ? -1.0 : it->second;
}

}  // extern "C"
