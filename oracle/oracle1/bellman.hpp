// Oracle-1: C++ restatement of bellman's Groth16 hot path (TEST INFRASTRUCTURE
// ONLY -- see field.hpp).  Templated on an engine so the DummyEngine
// instantiation reproduces the reference's known-answer test
// (groth16/src/tests/mod.rs:91-373) and the BLS12-381 instantiation serves as
// (i) the parity oracle for the CUDA path and (ii) the timed "restated bellman
// CPU path" baseline (same algorithmic choices as the rayon code: c = ceil(ln n),
// one task per window, all MSMs of a proof in flight, 2^k-way split FFT).
//
// All file:line citations are relative to /root/reference.
#pragma once
#include <cmath>
#include <condition_variable>
#include <functional>
#include <future>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <thread>
#include <vector>

#include "curve.hpp"

namespace o1 {

// ---------------------------------------------------------------------------
// src/multicore.rs -- Worker over a global pool (rayon stand-in)
// ---------------------------------------------------------------------------
class Pool {
  public:
    explicit Pool(unsigned n) : n_(n ? n : 1) {
        for (unsigned i = 0; i < n_; i++) ths_.emplace_back([this] { run(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : ths_) t.join();
    }
    unsigned size() const { return n_; }
    template <class F>
    auto spawn(F f) -> std::future<decltype(f())> {
        auto task = std::make_shared<std::packaged_task<decltype(f())()>>(std::move(f));
        auto fut = task->get_future();
        { std::lock_guard<std::mutex> g(mu_); q_.push([task] { (*task)(); }); }
        cv_.notify_one();
        return fut;
    }
  private:
    void run() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                job = std::move(q_.front());
                q_.pop();
            }
            job();
        }
    }
    unsigned n_;
    std::vector<std::thread> ths_;
    std::queue<std::function<void()>> q_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

struct Worker {
    Pool* pool;
    unsigned num_threads() const { return pool->size(); }
    uint32_t log_num_threads() const {           // multicore.rs:29-31,120-130
        uint32_t pow = 0;
        while ((1u << (pow + 1)) <= num_threads()) pow++;
        return pow;
    }
    // Worker::scope (multicore.rs:78-91): chunk = elements/threads (>=1); the body
    // gets (chunk_index, begin, end) for every chunk, all run on the pool, joined.
    template <class F>
    void scope_chunks(size_t elements, F body) const {
        size_t nt = num_threads();
        size_t chunk = elements < nt ? 1 : elements / nt;
        std::vector<std::future<void>> futs;
        size_t idx = 0;
        for (size_t b = 0; b < elements; b += chunk, idx++) {
            size_t e = std::min(elements, b + chunk);
            futs.push_back(pool->spawn([=] { body(idx, b, e); }));
        }
        for (auto& f : futs) f.get();
    }
};

// ---------------------------------------------------------------------------
// errors (src/lib.rs:304-319)
// ---------------------------------------------------------------------------
enum Err { OK = 0, ERR_POLY_DEGREE_TOO_LARGE = 1, ERR_UNEXPECTED_IDENTITY = 2, ERR_IO_EOF = 3,
           ERR_UNCONSTRAINED_VARIABLE = 4, ERR_DENSITY_MISMATCH = 5 };
struct SynthesisError { Err code; };

// ---------------------------------------------------------------------------
// src/domain.rs
// ---------------------------------------------------------------------------
inline uint32_t bitreverse(uint32_t n, uint32_t l) {        // domain.rs:273-280
    uint32_t r = 0;
    for (uint32_t i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}

template <class S>
void serial_fft(S* a, size_t len, const S& omega, uint32_t log_n) {   // domain.rs:272-314
    uint32_t n = (uint32_t)len;
    if (n != (1u << log_n)) throw std::logic_error("serial_fft size");
    for (uint32_t k = 0; k < n; k++) {
        uint32_t rk = bitreverse(k, log_n);
        if (k < rk) std::swap(a[rk], a[k]);
    }
    uint32_t m = 1;
    for (uint32_t s = 0; s < log_n; s++) {
        S w_m = omega.pow((uint64_t)(n / (2 * m)));
        for (uint32_t k = 0; k < n; k += 2 * m) {
            S w = S::one();
            for (uint32_t j = 0; j < m; j++) {
                S t = a[k + j + m] * w;
                S tmp = a[k + j] - t;
                a[k + j + m] = tmp;
                a[k + j] = a[k + j] + t;
                w = w * w_m;
            }
        }
        m *= 2;
    }
}

template <class S>
void parallel_fft(std::vector<S>& a, const Worker& worker, const S& omega, uint32_t log_n,
                  uint32_t log_cpus) {                      // domain.rs:316-372
    size_t num_cpus = (size_t)1 << log_cpus;
    uint32_t log_new_n = log_n - log_cpus;
    std::vector<std::vector<S>> tmp(num_cpus, std::vector<S>((size_t)1 << log_new_n, S::zero()));
    S new_omega = omega.pow((uint64_t)num_cpus);
    {
        std::vector<std::future<void>> futs;
        for (size_t j = 0; j < num_cpus; j++) {
            futs.push_back(worker.pool->spawn([&, j] {
                S omega_j = omega.pow((uint64_t)j);
                S omega_step = omega.pow((uint64_t)j << log_new_n);
                S elt = S::one();
                std::vector<S>& t = tmp[j];
                for (size_t i = 0; i < t.size(); i++) {
                    for (size_t s = 0; s < num_cpus; s++) {
                        size_t idx = (i + (s << log_new_n)) % ((size_t)1 << log_n);
                        t[i] = t[i] + a[idx] * elt;
                        elt = elt * omega_step;
                    }
                    elt = elt * omega_j;
                }
                serial_fft(t.data(), t.size(), new_omega, log_new_n);
            }));
        }
        for (auto& f : futs) f.get();
    }
    size_t mask = ((size_t)1 << log_cpus) - 1;
    worker.scope_chunks(a.size(), [&](size_t, size_t b, size_t e) {
        for (size_t idx = b; idx < e; idx++) a[idx] = tmp[idx & mask][idx >> log_cpus];
    });
}

template <class S>
void best_fft(std::vector<S>& a, const Worker& worker, const S& omega, uint32_t log_n) {  // :261-269
    uint32_t log_cpus = worker.log_num_threads();
    if (log_n <= log_cpus) serial_fft(a.data(), a.size(), omega, log_n);
    else parallel_fft(a, worker, omega, log_n, log_cpus);
}

template <class S>
struct EvaluationDomain {
    std::vector<S> coeffs;
    uint32_t exp;
    S omega, omegainv, geninv, minv;

    static EvaluationDomain from_coeffs(std::vector<S> c) {    // domain.rs:47-79
        size_t m = 1;
        uint32_t exp = 0;
        while (m < c.size()) {
            m *= 2;
            exp += 1;
            if (exp >= S::S) throw SynthesisError{ERR_POLY_DEGREE_TOO_LARGE};
        }
        S omega = S::root_of_unity();
        for (uint32_t i = exp; i < S::S; i++) omega = omega.square();
        c.resize(m, S::zero());
        EvaluationDomain d;
        d.coeffs = std::move(c);
        d.exp = exp;
        d.omega = omega;
        d.omegainv = omega.inv();
        d.geninv = S::generator().inv();
        d.minv = S::from_u64((uint64_t)m).inv();
        return d;
    }
    void fft(const Worker& w) { best_fft(coeffs, w, omega, exp); }                 // :81-83
    void ifft(const Worker& w) {                                                   // :85-99
        best_fft(coeffs, w, omegainv, exp);
        S mi = minv;
        w.scope_chunks(coeffs.size(), [&](size_t, size_t b, size_t e) {
            for (size_t i = b; i < e; i++) coeffs[i] = coeffs[i] * mi;
        });
    }
    void distribute_powers(const Worker& w, const S& g) {                          // :101-113
        w.scope_chunks(coeffs.size(), [&](size_t, size_t b, size_t e) {
            S u = g.pow((uint64_t)b);
            for (size_t i = b; i < e; i++) { coeffs[i] = coeffs[i] * u; u = u * g; }
        });
    }
    void coset_fft(const Worker& w) { distribute_powers(w, S::generator()); fft(w); }   // :115-118
    void icoset_fft(const Worker& w) { S gi = geninv; ifft(w); distribute_powers(w, gi); }  // :120-125
    S z(const S& tau) const { return tau.pow((uint64_t)coeffs.size()) - S::one(); }     // :129-134
    void divide_by_z_on_coset(const Worker& w) {                                   // :139-151
        S i = z(S::generator()).inv();
        w.scope_chunks(coeffs.size(), [&](size_t, size_t b, size_t e) {
            for (size_t k = b; k < e; k++) coeffs[k] = coeffs[k] * i;
        });
    }
    void mul_assign(const Worker& w, const EvaluationDomain& o) {                  // :154-170
        if (coeffs.size() != o.coeffs.size()) throw std::logic_error("mul_assign size");
        w.scope_chunks(coeffs.size(), [&](size_t, size_t b, size_t e) {
            for (size_t k = b; k < e; k++) coeffs[k] = coeffs[k] * o.coeffs[k];
        });
    }
    void sub_assign(const Worker& w, const EvaluationDomain& o) {                  // :173-189
        if (coeffs.size() != o.coeffs.size()) throw std::logic_error("sub_assign size");
        w.scope_chunks(coeffs.size(), [&](size_t, size_t b, size_t e) {
            for (size_t k = b; k < e; k++) coeffs[k] = coeffs[k] - o.coeffs[k];
        });
    }
};

// ---------------------------------------------------------------------------
// src/multiexp.rs
// ---------------------------------------------------------------------------
inline uint32_t window_size(size_t n) {                    // multiexp.rs:318-322
    return n < 32 ? 3u : (uint32_t)std::ceil(std::log((double)(uint32_t)n));
}

// Exponent<F> (multiexp.rs:159-182): 0 = Zero, 1 = One, 2 = Bits
struct Exponent {
    uint8_t kind;
    std::array<uint64_t, 4> bits;
};
template <class S>
Exponent to_exponent(const S& s) {
    Exponent e;
    e.bits = s.to_bits();
    if (s.is_zero()) e.kind = 0;
    else if (s == S::one()) e.kind = 1;
    else e.kind = 2;
    return e;
}
inline uint64_t exp_digit(const Exponent& e, uint32_t chunk, uint32_t c) {   // chunks(), :190-208
    uint32_t lo = chunk * c;
    uint64_t d = 0;
    for (uint32_t i = 0; i < c; i++) {
        uint32_t bit = lo + i;
        if (bit >= 256) break;
        d |= ((e.bits[bit / 64] >> (bit % 64)) & 1) << i;
    }
    return d;
}

// density: nullptr = FullDensity; else one byte per exponent (0/1)
template <class G, class GA>
struct MultiexpHandle {
    std::vector<std::future<std::pair<Err, G>>> parts;
    uint32_t c;
    std::pair<Err, G> wait() {                             // fold, multiexp.rs:295-300
        std::vector<std::pair<Err, G>> got;
        for (auto& f : parts) got.push_back(f.get());
        G acc = G::identity();
        for (size_t i = got.size(); i-- > 0;) {
            if (got[i].first != OK) return {got[i].first, G::identity()};
            for (uint32_t k = 0; k < c; k++) acc = acc.dbl();
            acc = acc.add(got[i].second);
        }
        return {OK, acc};
    }
};

template <class G, class GA>
MultiexpHandle<G, GA> multiexp(const Worker& worker, const GA* bases, size_t n_bases, size_t offset,
                               const uint8_t* density, size_t density_len,
                               const Exponent* exps, size_t n, uint32_t num_bits) {
    uint32_t c = window_size(n);                           // multiexp.rs:318-322
    if (density && density_len != n) throw SynthesisError{ERR_DENSITY_MISMATCH};   // :324-329 (assert)
    MultiexpHandle<G, GA> h;
    h.c = c;
    uint32_t chunk = 0;
    for (uint32_t lo = 0; lo < num_bits; lo += c, chunk++) {    // :288-293
        h.parts.push_back(worker.pool->spawn([=]() -> std::pair<Err, G> {
            G acc = G::identity();                         // :230
            size_t pos = offset;                           // Source (multiexp.rs:45-86)
            std::vector<G> buckets(((size_t)1 << c) - 1, G::identity());   // :236
            bool handle_trivial = chunk == 0;
            auto next = [&](const GA*& out) -> Err {       // :54-71
                if (n_bases <= pos) return ERR_IO_EOF;
                if (bases[pos].is_identity()) return ERR_UNEXPECTED_IDENTITY;
                out = &bases[pos++];
                return OK;
            };
            auto skip = [&]() -> Err {                     // :73-85
                if (n_bases <= pos) return ERR_IO_EOF;
                pos += 1;
                return OK;
            };
            for (size_t i = 0; i < n; i++) {               // :242-265
                if (density && !density[i]) continue;
                const Exponent& e = exps[i];
                Err er = OK;
                const GA* b = nullptr;
                if (e.kind == 0) er = skip();
                else if (e.kind == 1) {
                    if (handle_trivial) { er = next(b); if (er == OK) acc = acc.add_mixed(*b); }
                    else er = skip();
                } else {
                    uint64_t d = exp_digit(e, chunk, c);
                    if (d != 0) { er = next(b); if (er == OK) buckets[d - 1] = buckets[d - 1].add_mixed(*b); }
                    else er = skip();
                }
                if (er != OK) return {er, G::identity()};
            }
            G running = G::identity();                     // :271-275
            for (size_t k = buckets.size(); k-- > 0;) {
                running = running.add(buckets[k]);
                acc = acc.add(running);
            }
            return {OK, acc};
        }));
    }
    return h;
}

// ---------------------------------------------------------------------------
// Circuits as explicit R1CS.  The Circuit/ConstraintSystem trait surface
// (src/lib.rs) is unchanged CPU code upstream of the hot path; the oracle only
// needs what ProvingAssignment / KeypairAssembly record.  A linear combination
// is a list of (is_input, index, coeff) in push order (src/lib.rs:190-299).
// ---------------------------------------------------------------------------
template <class S>
struct Term { bool is_input; size_t idx; S coeff; };
template <class S>
struct R1CS {
    size_t num_inputs = 0, num_aux = 0;                    // num_inputs counts the ONE input
    std::vector<std::vector<Term<S>>> A, B, C;             // per constraint
    std::vector<S> input_assignment, aux_assignment;       // witness (may be empty for keygen)
    // builder helpers mirroring alloc/alloc_input/enforce call order
    size_t alloc(const S& v) { aux_assignment.push_back(v); return num_aux++; }
    size_t alloc_input(const S& v) { input_assignment.push_back(v); return num_inputs++; }
    void enforce(std::vector<Term<S>> a, std::vector<Term<S>> b, std::vector<Term<S>> c) {
        A.push_back(std::move(a)); B.push_back(std::move(b)); C.push_back(std::move(c));
    }
    // prover.rs:208-215 / generator.rs:195-202: x_i * 0 = 0 for every input
    void add_input_constraints() {
        for (size_t i = 0; i < num_inputs; i++) enforce({{true, i, S::one()}}, {}, {});
    }
};

// ProvingAssignment bookkeeping (prover.rs:19-55,105-145)
template <class S>
struct WitnessEval {
    std::vector<S> a, b, c;
    std::vector<uint8_t> a_aux_density, b_input_density, b_aux_density;
};
template <class S>
WitnessEval<S> eval_witness(const R1CS<S>& cs) {
    WitnessEval<S> w;
    w.a_aux_density.assign(cs.num_aux, 0);
    w.b_input_density.assign(cs.num_inputs, 0);
    w.b_aux_density.assign(cs.num_aux, 0);
    auto ev = [&](const std::vector<Term<S>>& lc, uint8_t* in_d, uint8_t* aux_d) {
        S acc = S::zero();
        for (const auto& t : lc) {
            if (t.coeff.is_zero()) continue;               // prover.rs:31
            S tmp;
            if (t.is_input) { tmp = cs.input_assignment[t.idx]; if (in_d) in_d[t.idx] = 1; }
            else { tmp = cs.aux_assignment[t.idx]; if (aux_d) aux_d[t.idx] = 1; }
            if (!(t.coeff == S::one())) tmp = tmp * t.coeff;
            acc = acc + tmp;
        }
        return acc;
    };
    for (size_t k = 0; k < cs.A.size(); k++) {
        w.a.push_back(ev(cs.A[k], nullptr, w.a_aux_density.data()));
        w.b.push_back(ev(cs.B[k], w.b_input_density.data(), w.b_aux_density.data()));
        w.c.push_back(ev(cs.C[k], nullptr, nullptr));
    }
    return w;
}

// ---------------------------------------------------------------------------
// groth16 Parameters (groth16/src/lib.rs:103-131,222-244)
// ---------------------------------------------------------------------------
template <class E>
struct Parameters {
    typename E::G1A alpha_g1, beta_g1, delta_g1;
    typename E::G2A beta_g2, gamma_g2, delta_g2;
    std::vector<typename E::G1A> ic, h, l, a, b_g1;
    std::vector<typename E::G2A> b_g2;
};

// QAP evaluation at tau: the scalars k such that CRS element = [k]G.
// (generator.rs:249-268 powers of tau + t(tau)/delta, :300 Lagrange ifft,
//  :376-415 eval_at_tau and the ext = (beta*a + alpha*b + c)/inv combination)
template <class S>
struct CrsScalars {
    size_t m;
    std::vector<S> h;                          // tau^i * t(tau)/delta, i < m-1
    std::vector<S> a_in, b_in, ext_in;         // per input
    std::vector<S> a_aux, b_aux, ext_aux;      // per aux
};
template <class S>
CrsScalars<S> crs_scalars(const R1CS<S>& cs, const Worker& worker, const S& alpha, const S& beta,
                          const S& gamma, const S& delta, const S& tau) {
    CrsScalars<S> out;
    auto dom = EvaluationDomain<S>::from_coeffs(std::vector<S>(cs.A.size(), S::zero()));
    size_t m = dom.coeffs.size();
    out.m = m;
    if (gamma.is_zero() || delta.is_zero()) throw SynthesisError{ERR_UNEXPECTED_IDENTITY};
    S gamma_inverse = gamma.inv(), delta_inverse = delta.inv();
    worker.scope_chunks(m, [&](size_t, size_t b, size_t e) {
        S cur = tau.pow((uint64_t)b);
        for (size_t i = b; i < e; i++) { dom.coeffs[i] = cur; cur = cur * tau; }
    });
    S coeff = dom.z(tau) * delta_inverse;
    out.h.resize(m - 1);
    for (size_t i = 0; i + 1 < m; i++) out.h[i] = dom.coeffs[i] * coeff;
    dom.ifft(worker);
    const std::vector<S>& lag = dom.coeffs;
    // transpose R1CS rows into per-variable sums (KeypairAssembly, generator.rs:43-155)
    std::vector<S> at_in(cs.num_inputs, S::zero()), bt_in = at_in, ct_in = at_in;
    std::vector<S> at_aux(cs.num_aux, S::zero()), bt_aux = at_aux, ct_aux = at_aux;
    auto acc = [&](const std::vector<std::vector<Term<S>>>& M, std::vector<S>& in, std::vector<S>& aux) {
        for (size_t k = 0; k < M.size(); k++)
            for (const auto& t : M[k]) {
                S& dst = t.is_input ? in[t.idx] : aux[t.idx];
                dst = dst + lag[k] * t.coeff;
            }
    };
    acc(cs.A, at_in, at_aux);
    acc(cs.B, bt_in, bt_aux);
    acc(cs.C, ct_in, ct_aux);
    auto ext = [&](const std::vector<S>& at, const std::vector<S>& bt, const std::vector<S>& ct,
                   const S& inv, std::vector<S>& o) {
        o.resize(at.size());
        for (size_t i = 0; i < at.size(); i++) o[i] = (at[i] * beta + bt[i] * alpha + ct[i]) * inv;
    };
    ext(at_in, bt_in, ct_in, gamma_inverse, out.ext_in);
    ext(at_aux, bt_aux, ct_aux, delta_inverse, out.ext_aux);
    out.a_in = std::move(at_in); out.b_in = std::move(bt_in);
    out.a_aux = std::move(at_aux); out.b_aux = std::move(bt_aux);
    return out;
}

// generate_parameters (generator.rs:159-507).  `mul1`/`mul2` = [k]g1 / [k]g2
// (the reference uses wNAF tables; the group elements are identical).
template <class E, class Mul1, class Mul2>
Parameters<E> generate_parameters(const R1CS<typename E::Fr>& cs_in, const Worker& worker,
                                  Mul1 mul1, Mul2 mul2, const typename E::Fr& alpha,
                                  const typename E::Fr& beta, const typename E::Fr& gamma,
                                  const typename E::Fr& delta, const typename E::Fr& tau) {
    typedef typename E::Fr S;
    typedef typename E::G1 G1;
    typedef typename E::G2 G2;
    R1CS<S> cs = cs_in;
    cs.add_input_constraints();
    CrsScalars<S> k = crs_scalars(cs, worker, alpha, beta, gamma, delta, tau);
    Parameters<E> p;
    auto map1 = [&](const std::vector<S>& ks, bool skip_zero) {
        std::vector<G1> proj(ks.size());
        worker.scope_chunks(ks.size(), [&](size_t, size_t b, size_t e) {
            for (size_t i = b; i < e; i++)
                proj[i] = (skip_zero && ks[i].is_zero()) ? G1::identity() : mul1(ks[i]);
        });
        std::vector<typename E::G1A> out(ks.size());
        E::batch_normalize1(proj, out.data());
        return out;
    };
    auto map2 = [&](const std::vector<S>& ks) {
        std::vector<G2> proj(ks.size());
        worker.scope_chunks(ks.size(), [&](size_t, size_t b, size_t e) {
            for (size_t i = b; i < e; i++) proj[i] = ks[i].is_zero() ? G2::identity() : mul2(ks[i]);
        });
        std::vector<typename E::G2A> out(ks.size());
        E::batch_normalize2(proj, out.data());
        return out;
    };
    p.h = map1(k.h, false);
    auto a_in = map1(k.a_in, true), a_aux = map1(k.a_aux, true);
    auto b1_in = map1(k.b_in, true), b1_aux = map1(k.b_aux, true);
    auto b2_in = map2(k.b_in), b2_aux = map2(k.b_aux);
    p.ic = map1(k.ext_in, false);
    p.l = map1(k.ext_aux, false);
    for (const auto& e : p.l)                              // generator.rs:466-470
        if (e.is_identity()) throw SynthesisError{ERR_UNCONSTRAINED_VARIABLE};
    auto filt = [](auto& dst, const auto& x, const auto& y) {   // generator.rs:491-505
        for (const auto& e : x) if (!e.is_identity()) dst.push_back(e);
        for (const auto& e : y) if (!e.is_identity()) dst.push_back(e);
    };
    filt(p.a, a_in, a_aux);
    filt(p.b_g1, b1_in, b1_aux);
    filt(p.b_g2, b2_in, b2_aux);
    p.alpha_g1 = mul1(alpha).to_affine(); p.beta_g1 = mul1(beta).to_affine();
    p.delta_g1 = mul1(delta).to_affine();
    p.beta_g2 = mul2(beta).to_affine(); p.gamma_g2 = mul2(gamma).to_affine();
    p.delta_g2 = mul2(delta).to_affine();
    return p;
}

// ---------------------------------------------------------------------------
// groth16/src/prover.rs
// ---------------------------------------------------------------------------
// H polynomial coefficients (prover.rs:221-242), m-1 of them
template <class S>
std::vector<S> h_coefficients(const Worker& worker, std::vector<S> av, std::vector<S> bv,
                              std::vector<S> cv) {
    auto a = EvaluationDomain<S>::from_coeffs(std::move(av));
    auto b = EvaluationDomain<S>::from_coeffs(std::move(bv));
    auto c = EvaluationDomain<S>::from_coeffs(std::move(cv));
    a.ifft(worker); a.coset_fft(worker);
    b.ifft(worker); b.coset_fft(worker);
    c.ifft(worker); c.coset_fft(worker);
    a.mul_assign(worker, b);
    a.sub_assign(worker, c);
    a.divide_by_z_on_coset(worker);
    a.icoset_fft(worker);
    a.coeffs.pop_back();                                    // prover.rs:238-240
    return std::move(a.coeffs);
}

template <class E>
struct Proof { typename E::G1A a; typename E::G2A b; typename E::G1A c; };

template <class E>
struct ProofDetails { std::vector<typename E::Fr> h_coeffs; };

// create_proof after synthesis (prover.rs:208-360).  `cs` carries the witness.
template <class E>
Proof<E> create_proof(const R1CS<typename E::Fr>& cs_in, const Parameters<E>& params,
                      const Worker& worker, const typename E::Fr& r, const typename E::Fr& s,
                      ProofDetails<E>* details = nullptr) {
    typedef typename E::Fr S;
    typedef typename E::G1 G1;
    typedef typename E::G2 G2;
    typedef typename E::G1A G1A;
    typedef typename E::G2A G2A;
    R1CS<S> cs = cs_in;
    cs.add_input_constraints();                                            // :208-215
    WitnessEval<S> w = eval_witness(cs);
    std::vector<S> hco = h_coefficients(worker, w.a, w.b, w.c);            // :221-240
    if (details) details->h_coeffs = hco;
    auto to_exps = [](const std::vector<S>& v) {                           // :242,248-261
        std::vector<Exponent> e(v.size());
        for (size_t i = 0; i < v.size(); i++) e[i] = to_exponent(v[i]);
        return e;
    };
    std::vector<Exponent> h_e = to_exps(hco), in_e = to_exps(cs.input_assignment),
                          aux_e = to_exps(cs.aux_assignment);
    const uint32_t NB = S::NUM_BITS;
    auto h = multiexp<G1, G1A>(worker, params.h.data(), params.h.size(), 0, nullptr, 0,
                               h_e.data(), h_e.size(), NB);                // :244
    auto l = multiexp<G1, G1A>(worker, params.l.data(), params.l.size(), 0, nullptr, 0,
                               aux_e.data(), aux_e.size(), NB);            // :263-268
    auto a_inputs = multiexp<G1, G1A>(worker, params.a.data(), params.a.size(), 0, nullptr, 0,
                                      in_e.data(), in_e.size(), NB);       // :275-280
    auto a_aux = multiexp<G1, G1A>(worker, params.a.data(), params.a.size(), in_e.size(),
                                   w.a_aux_density.data(), w.a_aux_density.size(),
                                   aux_e.data(), aux_e.size(), NB);        // :281-286
    size_t b_in_total = 0;
    for (uint8_t d : w.b_input_density) b_in_total += d;                   // :288-291
    auto b_g1_inputs = multiexp<G1, G1A>(worker, params.b_g1.data(), params.b_g1.size(), 0,
                                         w.b_input_density.data(), w.b_input_density.size(),
                                         in_e.data(), in_e.size(), NB);    // :296-301
    auto b_g1_aux = multiexp<G1, G1A>(worker, params.b_g1.data(), params.b_g1.size(), b_in_total,
                                      w.b_aux_density.data(), w.b_aux_density.size(),
                                      aux_e.data(), aux_e.size(), NB);     // :302-307
    auto b_g2_inputs = multiexp<G2, G2A>(worker, params.b_g2.data(), params.b_g2.size(), 0,
                                         w.b_input_density.data(), w.b_input_density.size(),
                                         in_e.data(), in_e.size(), NB);    // :312-317
    auto b_g2_aux = multiexp<G2, G2A>(worker, params.b_g2.data(), params.b_g2.size(), b_in_total,
                                      w.b_aux_density.data(), w.b_aux_density.size(),
                                      aux_e.data(), aux_e.size(), NB);     // :318
    if (params.delta_g1.is_identity() || params.delta_g2.is_identity())    // :320-324
        throw SynthesisError{ERR_UNEXPECTED_IDENTITY};
    G1 g_a = G1::from_affine(params.delta_g1).mul(r).add_mixed(params.alpha_g1);   // :326-327
    G2 g_b = G2::from_affine(params.delta_g2).mul(s).add_mixed(params.beta_g2);    // :328-329
    S rs = r * s;
    G1 g_c = G1::from_affine(params.delta_g1).mul(rs);                     // :335-337
    g_c = g_c.add(G1::from_affine(params.alpha_g1).mul(s));
    g_c = g_c.add(G1::from_affine(params.beta_g1).mul(r));
    auto get = [](auto& hnd) {
        auto res = hnd.wait();
        if (res.first != OK) throw SynthesisError{res.first};
        return res.second;
    };
    G1 a_answer = get(a_inputs);                                           // :339-343
    a_answer = a_answer.add(get(a_aux));
    g_a = g_a.add(a_answer);
    g_c = g_c.add(a_answer.mul(s));
    G1 b1_answer = get(b_g1_inputs);                                       // :345-354
    b1_answer = b1_answer.add(get(b_g1_aux));
    G2 b2_answer = get(b_g2_inputs);
    b2_answer = b2_answer.add(get(b_g2_aux));
    g_b = g_b.add(b2_answer);
    g_c = g_c.add(b1_answer.mul(r));
    g_c = g_c.add(get(h));
    g_c = g_c.add(get(l));
    return {g_a.to_affine(), g_b.to_affine(), g_c.to_affine()};            // :356-360
}

// ---------------------------------------------------------------------------
// Circuits
// ---------------------------------------------------------------------------
// MiMC (groth16/tests/common/mod.rs:48-129); rounds = 322 in the reference test,
// R = 524287 for the synthetic 2^20-constraint workload.
// with_witness=false leaves assignments as zeros (keygen only needs the shape).
template <class S>
R1CS<S> mimc_circuit(const S& xl0, const S& xr0, const std::vector<S>& constants) {
    R1CS<S> cs;
    cs.alloc_input(S::one());
    S xl_v = xl0, xr_v = xr0;
    size_t xl = cs.alloc(xl_v), xr = cs.alloc(xr_v);
    bool xl_is_input = false;
    size_t rounds = constants.size();
    for (size_t i = 0; i < rounds; i++) {
        const S& ci = constants[i];
        S t = xl_v + ci;
        S tmp_v = t.square();
        size_t tmp = cs.alloc(tmp_v);
        cs.enforce({{xl_is_input, xl, S::one()}, {true, 0, ci}},
                   {{xl_is_input, xl, S::one()}, {true, 0, ci}}, {{false, tmp, S::one()}});
        S new_v = t * tmp_v + xr_v;
        bool new_is_input = (i == rounds - 1);
        size_t nw = new_is_input ? cs.alloc_input(new_v) : cs.alloc(new_v);
        cs.enforce({{false, tmp, S::one()}}, {{xl_is_input, xl, S::one()}, {true, 0, ci}},
                   {{new_is_input, nw, S::one()}, {false, xr, S::zero() - S::one()}});
        // note: xr is never an input variable (it trails xl by one round)
        xr = xl; xr_v = xl_v;
        xl = nw; xl_v = new_v; xl_is_input = new_is_input;
    }
    return cs;
}

}  // namespace o1
