// Oracle-1 C API for the Python test harness (ctypes).  TEST INFRASTRUCTURE ONLY:
// loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs; never by the product path.
//
// Array formats (shared with the product's C-ABI so buffers can be compared
// byte-for-byte):
//   Fr      : 4 x u64 little-endian limbs, Montgomery form (R = 2^256) unless a
//             function says "canonical"
//   G1 aff. : 12 x u64 = x | y, Montgomery Fp limbs; all-zero = identity
//   G2 aff. : 24 x u64 = x.c0 | x.c1 | y.c0 | y.c1; all-zero = identity
#include <cstdio>
#include <memory>

#include "bellman.hpp"
#include "dummy_engine.hpp"

using namespace o1;

namespace {

struct Bls12 {
    typedef o1::Fr Fr;
    typedef o1::G1 G1;
    typedef o1::G1Affine G1A;
    typedef o1::G2 G2;
    typedef o1::G2Affine G2A;
    static void batch_normalize1(const std::vector<G1>& in, G1A* out) { batch_to_affine(in, out); }
    static void batch_normalize2(const std::vector<G2>& in, G2A* out) { batch_to_affine(in, out); }
};

std::unique_ptr<Pool> g_pool;
Worker worker() {
    if (!g_pool) g_pool.reset(new Pool(std::thread::hardware_concurrency()));
    return Worker{g_pool.get()};
}

inline Fr ld_fr(const uint64_t* p) { return Fr(Mont<FrParams>::from_raw(p)); }
inline void st_fr(uint64_t* p, const Fr& v) { std::memcpy(p, v.l, 32); }
inline Fp ld_fp(const uint64_t* p) { return Fp::from_raw(p); }
inline void st_fp(uint64_t* p, const Fp& v) { std::memcpy(p, v.l, 48); }

inline bool all_zero(const uint64_t* p, int n) { uint64_t a = 0; for (int i = 0; i < n; i++) a |= p[i]; return a == 0; }
inline G1Affine ld_g1(const uint64_t* p) {
    if (all_zero(p, 12)) return G1Affine::identity();
    return {ld_fp(p), ld_fp(p + 6), false};
}
inline void st_g1(uint64_t* p, const G1Affine& a) {
    if (a.inf) { std::memset(p, 0, 96); return; }
    st_fp(p, a.x); st_fp(p + 6, a.y);
}
inline G2Affine ld_g2(const uint64_t* p) {
    if (all_zero(p, 24)) return G2Affine::identity();
    return {{ld_fp(p), ld_fp(p + 6)}, {ld_fp(p + 12), ld_fp(p + 18)}, false};
}
inline void st_g2(uint64_t* p, const G2Affine& a) {
    if (a.inf) { std::memset(p, 0, 192); return; }
    st_fp(p, a.x.c0); st_fp(p + 6, a.x.c1); st_fp(p + 12, a.y.c0); st_fp(p + 18, a.y.c1);
}

// documented PRNG for synthetic inputs: splitmix64, rejection-sampled to [0, r)
struct SplitMix { uint64_t s; uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); } };
Fr random_fr(SplitMix& g) {
    for (;;) {
        uint64_t c[4] = {g.next(), g.next(), g.next(), g.next() & 0x7fffffffffffffffull};
        if (!Mont<FrParams>::geq_mod(c)) return Fr(Mont<FrParams>::from_canonical(c));
    }
}

const FixedBase<Fp>& fb1() { static FixedBase<Fp> t(g1_generator()); return t; }
const FixedBase<Fp2>& fb2() { static FixedBase<Fp2> t(g2_generator()); return t; }

struct MimcCase {
    R1CS<Fr> cs;                    // without the trailing input constraints
    std::vector<Fr> constants;
    Parameters<Bls12> params;
    bool have_params = false;
    Fr alpha, beta, gamma, delta, tau;
};

}  // namespace

extern "C" {

void o1_set_threads(int n) { g_pool.reset(new Pool(n > 0 ? n : std::thread::hardware_concurrency())); }
int o1_num_threads() { return (int)worker().num_threads(); }

// ---- field element-wise ops (n elements) ----------------------------------
void o1_fr_from_canonical(const uint64_t* in, uint64_t* out, size_t n) { for (size_t i = 0; i < n; i++) st_fr(out + 4 * i, Fr(Mont<FrParams>::from_canonical(in + 4 * i))); }
void o1_fr_to_canonical(const uint64_t* in, uint64_t* out, size_t n) { for (size_t i = 0; i < n; i++) ld_fr(in + 4 * i).to_canonical(out + 4 * i); }
void o1_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_fr(o + 4 * i, ld_fr(a + 4 * i) * ld_fr(b + 4 * i)); }
void o1_fr_add(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_fr(o + 4 * i, ld_fr(a + 4 * i) + ld_fr(b + 4 * i)); }
void o1_fr_sub(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_fr(o + 4 * i, ld_fr(a + 4 * i) - ld_fr(b + 4 * i)); }
void o1_fr_inv(const uint64_t* a, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_fr(o + 4 * i, ld_fr(a + 4 * i).inv()); }
void o1_fp_from_canonical(const uint64_t* in, uint64_t* out, size_t n) { for (size_t i = 0; i < n; i++) st_fp(out + 6 * i, Fp::from_canonical(in + 6 * i)); }
void o1_fp_to_canonical(const uint64_t* in, uint64_t* out, size_t n) { for (size_t i = 0; i < n; i++) ld_fp(in + 6 * i).to_canonical(out + 6 * i); }
void o1_fp_mul(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_fp(o + 6 * i, ld_fp(a + 6 * i) * ld_fp(b + 6 * i)); }
void o1_fp_add(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_fp(o + 6 * i, ld_fp(a + 6 * i) + ld_fp(b + 6 * i)); }
void o1_fp_sub(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_fp(o + 6 * i, ld_fp(a + 6 * i) - ld_fp(b + 6 * i)); }
void o1_fr_random(uint64_t seed, uint64_t* out, size_t n) { SplitMix g{seed}; for (size_t i = 0; i < n; i++) st_fr(out + 4 * i, random_fr(g)); }

// ---- curve ops ---------------------------------------------------------------
void o1_g1_generator(uint64_t* out) { st_g1(out, g1_generator().to_affine()); }
void o1_g2_generator(uint64_t* out) { st_g2(out, g2_generator().to_affine()); }
int o1_g1_on_curve(const uint64_t* p, size_t n) { for (size_t i = 0; i < n; i++) if (!g1_on_curve(ld_g1(p + 12 * i))) return 0; return 1; }
int o1_g2_on_curve(const uint64_t* p, size_t n) { for (size_t i = 0; i < n; i++) if (!g2_on_curve(ld_g2(p + 24 * i))) return 0; return 1; }
void o1_g1_add(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_g1(o + 12 * i, G1::from_affine(ld_g1(a + 12 * i)).add_mixed(ld_g1(b + 12 * i)).to_affine()); }
void o1_g2_add(const uint64_t* a, const uint64_t* b, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_g2(o + 24 * i, G2::from_affine(ld_g2(a + 24 * i)).add_mixed(ld_g2(b + 24 * i)).to_affine()); }
// out[i] = [k_i] base_i   (scalars Montgomery)
void o1_g1_mul(const uint64_t* bases, const uint64_t* k, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_g1(o + 12 * i, G1::from_affine(ld_g1(bases + 12 * i)).mul(ld_fr(k + 4 * i)).to_affine()); }
void o1_g2_mul(const uint64_t* bases, const uint64_t* k, uint64_t* o, size_t n) { for (size_t i = 0; i < n; i++) st_g2(o + 24 * i, G2::from_affine(ld_g2(bases + 24 * i)).mul(ld_fr(k + 4 * i)).to_affine()); }
// out[i] = [k_i] generator, threaded, fixed-base table
void o1_g1_fixed_mul(const uint64_t* k, uint64_t* o, size_t n) {
    const auto& t = fb1();
    worker().scope_chunks(n, [&](size_t, size_t b, size_t e) {
        std::vector<G1> proj; proj.reserve(e - b);
        for (size_t i = b; i < e; i++) proj.push_back(t.mul(ld_fr(k + 4 * i)));
        std::vector<G1Affine> aff(e - b); batch_to_affine(proj, aff.data());
        for (size_t i = b; i < e; i++) st_g1(o + 12 * i, aff[i - b]);
    });
}
void o1_g2_fixed_mul(const uint64_t* k, uint64_t* o, size_t n) {
    const auto& t = fb2();
    worker().scope_chunks(n, [&](size_t, size_t b, size_t e) {
        std::vector<G2> proj; proj.reserve(e - b);
        for (size_t i = b; i < e; i++) proj.push_back(t.mul(ld_fr(k + 4 * i)));
        std::vector<G2Affine> aff(e - b); batch_to_affine(proj, aff.data());
        for (size_t i = b; i < e; i++) st_g2(o + 24 * i, aff[i - b]);
    });
}
void o1_g1_compress(const uint64_t* p, uint8_t* out, size_t n) { for (size_t i = 0; i < n; i++) g1_compress(ld_g1(p + 12 * i), out + 48 * i); }
void o1_g2_compress(const uint64_t* p, uint8_t* out, size_t n) { for (size_t i = 0; i < n; i++) g2_compress(ld_g2(p + 24 * i), out + 96 * i); }

// ---- domain.rs ------------------------------------------------------------------
// mode: 0 fft, 1 ifft, 2 coset_fft, 3 icoset_fft (best_fft: split on the pool)
int o1_fft(uint64_t* data, uint32_t log_n, int mode) {
    size_t n = (size_t)1 << log_n;
    std::vector<Fr> v(n);
    for (size_t i = 0; i < n; i++) v[i] = ld_fr(data + 4 * i);
    try {
        auto d = EvaluationDomain<Fr>::from_coeffs(std::move(v));
        Worker w = worker();
        if (mode == 0) d.fft(w); else if (mode == 1) d.ifft(w); else if (mode == 2) d.coset_fft(w); else d.icoset_fft(w);
        for (size_t i = 0; i < n; i++) st_fr(data + 4 * i, d.coeffs[i]);
    } catch (SynthesisError& e) { return e.code; }
    return 0;
}
// serial_fft (domain.rs:272-314) with omega (inverse=0) or omega^-1, no scaling
void o1_serial_fft(uint64_t* data, uint32_t log_n, int inverse) {
    size_t n = (size_t)1 << log_n;
    std::vector<Fr> v(n);
    for (size_t i = 0; i < n; i++) v[i] = ld_fr(data + 4 * i);
    auto d = EvaluationDomain<Fr>::from_coeffs(std::vector<Fr>(n, Fr::zero()));
    serial_fft(v.data(), n, inverse ? d.omegainv : d.omega, log_n);
    for (size_t i = 0; i < n; i++) st_fr(data + 4 * i, v[i]);
}
// prover.rs:221-240: a,b,c (n_constraints each) -> out (m-1 coefficients); returns m, or -err
long o1_h_poly(const uint64_t* a, const uint64_t* b, const uint64_t* c, size_t n, uint64_t* out) {
    std::vector<Fr> av(n), bv(n), cv(n);
    for (size_t i = 0; i < n; i++) { av[i] = ld_fr(a + 4 * i); bv[i] = ld_fr(b + 4 * i); cv[i] = ld_fr(c + 4 * i); }
    try {
        auto h = h_coefficients(worker(), std::move(av), std::move(bv), std::move(cv));
        for (size_t i = 0; i < h.size(); i++) st_fr(out + 4 * i, h[i]);
        return (long)h.size() + 1;
    } catch (SynthesisError& e) { return -(long)e.code; }
}

// ---- multiexp.rs ----------------------------------------------------------------
uint32_t o1_window_size(size_t n) { return window_size(n); }
// density: NULL (FullDensity) or n bytes.  scalars Montgomery.  Returns Err code.
int o1_multiexp_g1(const uint64_t* bases, size_t n_bases, size_t offset, const uint8_t* density,
                   const uint64_t* scalars, size_t n, uint64_t* out) {
    std::vector<G1Affine> b(n_bases);
    for (size_t i = 0; i < n_bases; i++) b[i] = ld_g1(bases + 12 * i);
    std::vector<Exponent> e(n);
    for (size_t i = 0; i < n; i++) e[i] = to_exponent(ld_fr(scalars + 4 * i));
    auto h = multiexp<G1, G1Affine>(worker(), b.data(), n_bases, offset, density, n, e.data(), n, Fr::NUM_BITS);
    auto r = h.wait();
    if (r.first != OK) return r.first;
    st_g1(out, r.second.to_affine());
    return 0;
}
int o1_multiexp_g2(const uint64_t* bases, size_t n_bases, size_t offset, const uint8_t* density,
                   const uint64_t* scalars, size_t n, uint64_t* out) {
    std::vector<G2Affine> b(n_bases);
    for (size_t i = 0; i < n_bases; i++) b[i] = ld_g2(bases + 24 * i);
    std::vector<Exponent> e(n);
    for (size_t i = 0; i < n; i++) e[i] = to_exponent(ld_fr(scalars + 4 * i));
    auto h = multiexp<G2, G2Affine>(worker(), b.data(), n_bases, offset, density, n, e.data(), n, Fr::NUM_BITS);
    auto r = h.wait();
    if (r.first != OK) return r.first;
    st_g2(out, r.second.to_affine());
    return 0;
}
// naive_multiexp (multiexp.rs:336-349): sum [k_i] P_i
void o1_naive_multiexp_g1(const uint64_t* bases, const uint64_t* scalars, size_t n, uint64_t* out) {
    G1 acc = G1::identity();
    for (size_t i = 0; i < n; i++) acc = acc.add(G1::from_affine(ld_g1(bases + 12 * i)).mul(ld_fr(scalars + 4 * i)));
    st_g1(out, acc.to_affine());
}
void o1_naive_multiexp_g2(const uint64_t* bases, const uint64_t* scalars, size_t n, uint64_t* out) {
    G2 acc = G2::identity();
    for (size_t i = 0; i < n; i++) acc = acc.add(G2::from_affine(ld_g2(bases + 24 * i)).mul(ld_fr(scalars + 4 * i)));
    st_g2(out, acc.to_affine());
}

// ---- MiMC cases (groth16/tests/common/mod.rs; rounds=322 is the reference test,
//      rounds=524287 the synthetic 2^20-constraint workload) ---------------------------
void* o1_mimc_new(size_t rounds, uint64_t seed) {
    auto* mc = new MimcCase();
    SplitMix g{seed};
    mc->constants.resize(rounds);
    for (auto& c : mc->constants) c = random_fr(g);
    Fr xl = random_fr(g), xr = random_fr(g);
    mc->cs = mimc_circuit(xl, xr, mc->constants);
    return mc;
}
void o1_mimc_free(void* h) { delete (MimcCase*)h; }
// shape[0..6] = num_inputs, num_aux, num_constraints (incl. input constraints), m,
//               a_aux_density_total, b_input_density_total, b_aux_density_total
void o1_mimc_shape(void* h, uint64_t* shape) {
    auto* mc = (MimcCase*)h;
    R1CS<Fr> cs = mc->cs; cs.add_input_constraints();
    WitnessEval<Fr> w = eval_witness(cs);
    size_t m = 1; while (m < cs.A.size()) m *= 2;
    shape[0] = cs.num_inputs; shape[1] = cs.num_aux; shape[2] = cs.A.size(); shape[3] = m;
    uint64_t t = 0; for (auto d : w.a_aux_density) t += d; shape[4] = t;
    t = 0; for (auto d : w.b_input_density) t += d; shape[5] = t;
    t = 0; for (auto d : w.b_aux_density) t += d; shape[6] = t;
}
// what ProvingAssignment holds after synthesis (prover.rs:57-71): a,b,c evaluations,
// assignments, three density byte-maps
void o1_mimc_witness(void* h, uint64_t* a, uint64_t* b, uint64_t* c, uint64_t* inputs, uint64_t* aux,
                     uint8_t* a_aux_d, uint8_t* b_in_d, uint8_t* b_aux_d) {
    auto* mc = (MimcCase*)h;
    R1CS<Fr> cs = mc->cs; cs.add_input_constraints();
    WitnessEval<Fr> w = eval_witness(cs);
    for (size_t i = 0; i < w.a.size(); i++) { st_fr(a + 4 * i, w.a[i]); st_fr(b + 4 * i, w.b[i]); st_fr(c + 4 * i, w.c[i]); }
    for (size_t i = 0; i < cs.num_inputs; i++) st_fr(inputs + 4 * i, cs.input_assignment[i]);
    for (size_t i = 0; i < cs.num_aux; i++) st_fr(aux + 4 * i, cs.aux_assignment[i]);
    std::memcpy(a_aux_d, w.a_aux_density.data(), cs.num_aux);
    std::memcpy(b_in_d, w.b_input_density.data(), cs.num_inputs);
    std::memcpy(b_aux_d, w.b_aux_density.data(), cs.num_aux);
}
// toxic = alpha,beta,gamma,delta,tau (Montgomery Fr x5)
void o1_mimc_set_toxic(void* h, const uint64_t* toxic) {
    auto* mc = (MimcCase*)h;
    mc->alpha = ld_fr(toxic); mc->beta = ld_fr(toxic + 4); mc->gamma = ld_fr(toxic + 8);
    mc->delta = ld_fr(toxic + 12); mc->tau = ld_fr(toxic + 16);
}
// The dlogs of every CRS element w.r.t. the generators (generator.rs:249-268,300,376-415),
// so a GPU fixed-base kernel can manufacture a valid CRS at sizes the CPU cannot.
// hk: m-1; a_k,b_k,ext_k: num_inputs+num_aux each (inputs first; ext = ic | l)
void o1_mimc_crs_scalars(void* h, uint64_t* hk, uint64_t* a_k, uint64_t* b_k, uint64_t* ext_k) {
    auto* mc = (MimcCase*)h;
    R1CS<Fr> cs = mc->cs; cs.add_input_constraints();
    auto k = crs_scalars(cs, worker(), mc->alpha, mc->beta, mc->gamma, mc->delta, mc->tau);
    for (size_t i = 0; i < k.h.size(); i++) st_fr(hk + 4 * i, k.h[i]);
    size_t ni = cs.num_inputs;
    for (size_t i = 0; i < ni; i++) { st_fr(a_k + 4 * i, k.a_in[i]); st_fr(b_k + 4 * i, k.b_in[i]); st_fr(ext_k + 4 * i, k.ext_in[i]); }
    for (size_t i = 0; i < cs.num_aux; i++) { st_fr(a_k + 4 * (ni + i), k.a_aux[i]); st_fr(b_k + 4 * (ni + i), k.b_aux[i]); st_fr(ext_k + 4 * (ni + i), k.ext_aux[i]); }
}
// generate_parameters on the CPU (fixed-base tables); returns Err
int o1_mimc_generate(void* h) {
    auto* mc = (MimcCase*)h;
    try {
        const auto& t1 = fb1(); const auto& t2 = fb2();
        mc->params = generate_parameters<Bls12>(mc->cs, worker(),
            [&](const Fr& k) { return t1.mul(k); }, [&](const Fr& k) { return t2.mul(k); },
            mc->alpha, mc->beta, mc->gamma, mc->delta, mc->tau);
        mc->have_params = true;
    } catch (SynthesisError& e) { return e.code; }
    return 0;
}
// sizes[0..5] = |ic|, |h|, |l|, |a|, |b_g1|, |b_g2|
void o1_mimc_param_sizes(void* h, uint64_t* sizes) {
    auto& p = ((MimcCase*)h)->params;
    sizes[0] = p.ic.size(); sizes[1] = p.h.size(); sizes[2] = p.l.size(); sizes[3] = p.a.size(); sizes[4] = p.b_g1.size(); sizes[5] = p.b_g2.size();
}
// vk_g1 = alpha_g1|beta_g1|delta_g1 (3x12 u64); vk_g2 = beta_g2|gamma_g2|delta_g2 (3x24 u64)
void o1_mimc_export_params(void* h, uint64_t* vk_g1, uint64_t* vk_g2, uint64_t* ic, uint64_t* hq, uint64_t* l,
                           uint64_t* a, uint64_t* b_g1, uint64_t* b_g2) {
    auto& p = ((MimcCase*)h)->params;
    st_g1(vk_g1, p.alpha_g1); st_g1(vk_g1 + 12, p.beta_g1); st_g1(vk_g1 + 24, p.delta_g1);
    st_g2(vk_g2, p.beta_g2); st_g2(vk_g2 + 24, p.gamma_g2); st_g2(vk_g2 + 48, p.delta_g2);
    for (size_t i = 0; i < p.ic.size(); i++) st_g1(ic + 12 * i, p.ic[i]);
    for (size_t i = 0; i < p.h.size(); i++) st_g1(hq + 12 * i, p.h[i]);
    for (size_t i = 0; i < p.l.size(); i++) st_g1(l + 12 * i, p.l[i]);
    for (size_t i = 0; i < p.a.size(); i++) st_g1(a + 12 * i, p.a[i]);
    for (size_t i = 0; i < p.b_g1.size(); i++) st_g1(b_g1 + 12 * i, p.b_g1[i]);
    for (size_t i = 0; i < p.b_g2.size(); i++) st_g2(b_g2 + 24 * i, p.b_g2[i]);
}
// install externally produced parameters (e.g. manufactured on the GPU from crs_scalars)
void o1_mimc_import_params(void* h, const uint64_t* vk_g1, const uint64_t* vk_g2, const uint64_t* hq, size_t nh,
                           const uint64_t* l, size_t nl, const uint64_t* a, size_t na,
                           const uint64_t* b_g1, size_t nb1, const uint64_t* b_g2, size_t nb2) {
    auto* mc = (MimcCase*)h; auto& p = mc->params;
    p.alpha_g1 = ld_g1(vk_g1); p.beta_g1 = ld_g1(vk_g1 + 12); p.delta_g1 = ld_g1(vk_g1 + 24);
    p.beta_g2 = ld_g2(vk_g2); p.gamma_g2 = ld_g2(vk_g2 + 24); p.delta_g2 = ld_g2(vk_g2 + 48);
    p.h.resize(nh); for (size_t i = 0; i < nh; i++) p.h[i] = ld_g1(hq + 12 * i);
    p.l.resize(nl); for (size_t i = 0; i < nl; i++) p.l[i] = ld_g1(l + 12 * i);
    p.a.resize(na); for (size_t i = 0; i < na; i++) p.a[i] = ld_g1(a + 12 * i);
    p.b_g1.resize(nb1); for (size_t i = 0; i < nb1; i++) p.b_g1[i] = ld_g1(b_g1 + 12 * i);
    p.b_g2.resize(nb2); for (size_t i = 0; i < nb2; i++) p.b_g2[i] = ld_g2(b_g2 + 24 * i);
    mc->have_params = true;
}
// create_proof (prover.rs:182-361) + Proof::write (lib.rs:39-45).  r,s Montgomery.  Returns Err.
int o1_mimc_prove(void* h, const uint64_t* r, const uint64_t* s, uint8_t* proof192) {
    auto* mc = (MimcCase*)h;
    if (!mc->have_params) return -1;
    try {
        Proof<Bls12> pf = create_proof<Bls12>(mc->cs, mc->params, worker(), ld_fr(r), ld_fr(s));
        g1_compress(pf.a, proof192); g2_compress(pf.b, proof192 + 48); g1_compress(pf.c, proof192 + 144);
    } catch (SynthesisError& e) { return e.code; }
    return 0;
}
// Expected proof computed in the exponent from the toxic waste, the way
// groth16/src/tests/mod.rs:287-370 checks test_xordemo -- no MSM, no FFT:
//   a = alpha + A(tau) + r delta ; b = beta + B(tau) + s delta ;
//   c = s a + r b - r s delta + sum_aux w_i ext_i + (A(tau)B(tau) - C(tau))/delta
void o1_mimc_expected_proof(void* h, const uint64_t* r_, const uint64_t* s_, uint8_t* proof192) {
    auto* mc = (MimcCase*)h;
    R1CS<Fr> cs = mc->cs; cs.add_input_constraints();
    Fr r = ld_fr(r_), s = ld_fr(s_);
    // A(tau) = sum_k lag_k * <A_k, w>  (same Lagrange basis as the generator)
    auto dom = EvaluationDomain<Fr>::from_coeffs(std::vector<Fr>(cs.A.size(), Fr::zero()));
    size_t m = dom.coeffs.size();
    Fr cur = Fr::one();
    for (size_t i = 0; i < m; i++) { dom.coeffs[i] = cur; cur = cur * mc->tau; }
    Fr t_tau = dom.z(mc->tau);
    dom.ifft(worker());
    WitnessEval<Fr> w = eval_witness(cs);       // w.a[k] = <A_k, w> etc. (zero-coeff terms contribute 0 either way)
    Fr At = Fr::zero(), Bt = Fr::zero(), Ct = Fr::zero();
    for (size_t k = 0; k < w.a.size(); k++) { At = At + dom.coeffs[k] * w.a[k]; Bt = Bt + dom.coeffs[k] * w.b[k]; Ct = Ct + dom.coeffs[k] * w.c[k]; }
    auto ks = crs_scalars(cs, worker(), mc->alpha, mc->beta, mc->gamma, mc->delta, mc->tau);
    Fr lsum = Fr::zero();
    for (size_t i = 0; i < cs.num_aux; i++) lsum = lsum + ks.ext_aux[i] * cs.aux_assignment[i];
    Fr a = mc->alpha + At + r * mc->delta;
    Fr b = mc->beta + Bt + s * mc->delta;
    Fr c = s * a + r * b - r * s * mc->delta + lsum + (At * Bt - Ct) * mc->delta.inv();
    (void)t_tau;
    g1_compress(fb1().mul(a).to_affine(), proof192);
    g2_compress(fb2().mul(b).to_affine(), proof192 + 48);
    g1_compress(fb1().mul(c).to_affine(), proof192 + 144);
}

// ---- DummyEngine known-answer runs (groth16/src/tests/mod.rs) ---------------------
// out: [0..7) h, [7..9) l, [9..13) a, [13..15) b_g1, [15..17) b_g2, [17..19) ic,
//      [19..25) alpha_g1,beta_g1,beta_g2,gamma_g2,delta_g1,delta_g2, [25..28) proof a,b,c,
//      [28..35) H coefficients, [35..41) sizes h,l,a,b_g1,b_g2,ic
int o1_dummy_xordemo(uint32_t* out) {
    typedef DummyEngine E;
    auto F = [](uint32_t v) { return DFr{v}; };
    auto build = [&](bool witness, bool a, bool b) {       // tests/mod.rs:19-89
        R1CS<DFr> cs;
        cs.alloc_input(DFr::one());
        DFr m1 = DFr::zero() - DFr::one();
        size_t av = cs.alloc(F(witness && a ? 1 : 0));
        cs.enforce({{true, 0, DFr::one()}, {false, av, m1}}, {{false, av, DFr::one()}}, {});
        size_t bv = cs.alloc(F(witness && b ? 1 : 0));
        cs.enforce({{true, 0, DFr::one()}, {false, bv, m1}}, {{false, bv, DFr::one()}}, {});
        size_t cv = cs.alloc_input(F(witness && (a ^ b) ? 1 : 0));
        cs.enforce({{false, av, DFr::one()}, {false, av, DFr::one()}}, {{false, bv, DFr::one()}},
                   {{false, av, DFr::one()}, {false, bv, DFr::one()}, {true, cv, m1}});
        return cs;
    };
    Pool pool(2);
    Worker w{&pool};
    DFr alpha = F(48577), beta = F(22580), gamma = F(53332), delta = F(5481), tau = F(3673);   // :95-99
    DG g{DFr::one()};
    try {
        auto params = generate_parameters<E>(build(false, false, false), w,
            [&](const DFr& k) { return g.mul(k); }, [&](const DFr& k) { return g.mul(k); },
            alpha, beta, gamma, delta, tau);
        if (params.h.size() != 7 || params.l.size() != 2 || params.a.size() != 4 || params.b_g1.size() != 2 ||
            params.b_g2.size() != 2 || params.ic.size() != 2) return -2;
        for (int i = 0; i < 7; i++) out[i] = params.h[i].e.v;
        for (int i = 0; i < 2; i++) out[7 + i] = params.l[i].e.v;
        for (int i = 0; i < 4; i++) out[9 + i] = params.a[i].e.v;
        for (int i = 0; i < 2; i++) { out[13 + i] = params.b_g1[i].e.v; out[15 + i] = params.b_g2[i].e.v; out[17 + i] = params.ic[i].e.v; }
        out[19] = params.alpha_g1.e.v; out[20] = params.beta_g1.e.v; out[21] = params.beta_g2.e.v;
        out[22] = params.gamma_g2.e.v; out[23] = params.delta_g1.e.v; out[24] = params.delta_g2.e.v;
        ProofDetails<E> det;
        auto pf = create_proof<E>(build(true, true, false), params, w, F(27134), F(17146), &det);   // :274-285
        out[25] = pf.a.e.v; out[26] = pf.b.e.v; out[27] = pf.c.e.v;
        if (det.h_coeffs.size() != 7) return -3;
        for (int i = 0; i < 7; i++) out[28 + i] = det.h_coeffs[i].v;
    } catch (SynthesisError& e) { return e.code; }
    return 0;
}
// zero-coefficient regression (tests/mod.rs:375-440).  Returns 1 if the proof verifies
// under the toy pairing (verifier.rs:46-52 with pairing = product), 0 if not, <0 on error.
int o1_dummy_zero_coeff(int one_var) {
    typedef DummyEngine E;
    auto F = [](uint32_t v) { return DFr{v}; };
    R1CS<DFr> cs;
    cs.alloc_input(DFr::one());
    size_t a = cs.alloc(F(5)), b = cs.alloc(F(6)), c = cs.alloc(F(30));
    if (one_var) cs.enforce({{false, a, DFr::one()}}, {{true, 0, DFr::zero()}, {false, b, DFr::one()}}, {{false, c, DFr::one()}});
    else cs.enforce({{false, a, DFr::one()}}, {{false, a, DFr::zero()}, {false, b, DFr::one()}}, {{false, c, DFr::one()}});
    Pool pool(2);
    Worker w{&pool};
    DFr alpha = F(48577), beta = F(22580), gamma = F(53332), delta = F(5481), tau = F(3673);
    DG g{DFr::one()};
    try {
        auto pk = generate_parameters<E>(cs, w, [&](const DFr& k) { return g.mul(k); }, [&](const DFr& k) { return g.mul(k); },
                                         alpha, beta, gamma, delta, tau);
        auto pf = create_proof<E>(cs, pk, w, F(27134), F(17146));
        if (pk.ic.size() != 1) return -2;
        DFr acc = pk.ic[0].e;                               // verifier.rs:31-35 (no public inputs)
        DFr lhs = pf.a.e * pf.b.e;
        DFr rhs = pk.alpha_g1.e * pk.beta_g2.e + acc * pk.gamma_g2.e + pf.c.e * pk.delta_g2.e;
        return lhs == rhs ? 1 : 0;
    } catch (SynthesisError& e) { return -(int)e.code - 10; }
}

}  // extern "C"
