// bellman_b200: synthetic workload for the benchmark -- the MiMC chain of
// /root/reference/groth16/tests/common/mod.rs:48-129 run through what
// groth16's ProvingAssignment records during synthesis (groth16/src/prover.rs:73-145,
// 193-215).  Upstream this is user circuit code plus bookkeeping on the CPU, unchanged by this
// back-end; it lives here only so that bench.py can produce a valid witness of the
// benchmark's size without going through the test oracle.  Host code, product-side field
// arithmetic (mp.cuh host path).
#include "bb_internal.cuh"

using namespace bb;

namespace {

struct SplitMix {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

// uniform in [0, r): rejection sampling on 255-bit draws
Fr random_fr(SplitMix& g) {
    for (;;) {
        uint64_t c[4] = {g.next(), g.next(), g.next(), g.next() & 0x7fffffffffffffffull};
        Fr v;
        std::memcpy(v.l, c, 32);
        bool lt = false;
        for (int i = 7; i >= 0; i--) {
            if (v.l[i] < bbc::FR_MOD[i]) { lt = true; break; }
            if (v.l[i] > bbc::FR_MOD[i]) break;
        }
        if (lt) return fr_from_canonical(v);
    }
}

inline void put(uint64_t* dst, size_t i, const Fr& v) { std::memcpy(dst + 4 * i, v.l, 32); }
inline void set_bit(uint64_t* bits, size_t i) { bits[i >> 6] |= 1ull << (i & 63); }

}  // namespace

extern "C" {

// shape[0..6] = num_inputs, num_aux, num_constraints, m, a_aux_density_total,
//               b_input_density_total, b_aux_density_total
int bb_synth_mimc_shape(size_t rounds, uint64_t* shape) {
    if (!rounds || !shape) return BB_ERR_ARG;
    size_t n = 2 * rounds + 2, m = 1;
    while (m < n) m *= 2;
    shape[0] = 2; shape[1] = 2 * rounds + 1; shape[2] = n; shape[3] = m;
    shape[4] = 2 * rounds; shape[5] = 1; shape[6] = rounds;
    return BB_OK;
}

// a,b,c: n Fr; inputs: 2 Fr; aux: 2*rounds+1 Fr (all Montgomery); density bitmaps zeroed and
// filled here (LSB-first words).  Constants and preimage: splitmix64(seed), rejection sampled.
int bb_synth_mimc_witness(size_t rounds, uint64_t seed, uint64_t* a, uint64_t* b, uint64_t* c, uint64_t* inputs, uint64_t* aux,
                          uint64_t* a_aux_density, uint64_t* b_input_density, uint64_t* b_aux_density) {
    if (!rounds || !a || !b || !c || !inputs || !aux || !a_aux_density || !b_input_density || !b_aux_density) return BB_ERR_ARG;
    const size_t num_aux = 2 * rounds + 1;
    std::memset(a_aux_density, 0, ((num_aux + 63) / 64) * 8);
    std::memset(b_aux_density, 0, ((num_aux + 63) / 64) * 8);
    b_input_density[0] = 0;
    SplitMix g{seed};
    std::vector<Fr> constants(rounds);
    for (auto& k : constants) k = random_fr(g);
    Fr xl = random_fr(g), xr = random_fr(g);
    const Fr one = fr_one();
    put(inputs, 0, one);                                  // alloc_input(ONE), prover.rs:204
    size_t n_aux = 0, row = 0;
    put(aux, n_aux, xl); size_t xl_idx = n_aux++;          // "preimage xl"
    put(aux, n_aux, xr); n_aux++;                          // "preimage xr" (only ever appears in C)
    bool xl_is_input = false;
    for (size_t i = 0; i < rounds; i++) {
        const Fr& ci = constants[i];
        Fr t = xl + ci;                                    // <xl + Ci*ONE, w>
        Fr tmp = t.sqr();
        put(aux, n_aux, tmp); size_t tmp_idx = n_aux++;
        // tmp = (xL + Ci)^2 :  A = B = xl + Ci, C = tmp
        put(a, row, t); put(b, row, t); put(c, row, tmp); row++;
        if (!xl_is_input) { set_bit(a_aux_density, xl_idx); set_bit(b_aux_density, xl_idx); }
        if (!ci.is_zero()) set_bit(b_input_density, 0);    // eval skips zero coefficients, prover.rs:31
        Fr nw = t * tmp + xr;
        bool last = i + 1 == rounds;
        size_t nw_idx = 0;
        if (last) put(inputs, 1, nw);                      // "image" public input
        else { put(aux, n_aux, nw); nw_idx = n_aux++; }
        // new_xL = xR + tmp*(xL + Ci) :  A = tmp, B = xl + Ci, C = new_xl - xr
        put(a, row, tmp); put(b, row, t); put(c, row, nw - xr); row++;
        set_bit(a_aux_density, tmp_idx);
        xr = xl;
        xl = nw; xl_idx = nw_idx; xl_is_input = last;
    }
    // x_i * 0 = 0 for every input (prover.rs:208-215)
    Fr image;
    std::memcpy(image.l, inputs + 4, 32);
    put(a, row, one); put(b, row, Fr::zero()); put(c, row, Fr::zero()); row++;
    put(a, row, image); put(b, row, Fr::zero()); put(c, row, Fr::zero()); row++;
    return (n_aux == num_aux && row == 2 * rounds + 2) ? BB_OK : BB_ERR_ARG;
}

}  // extern "C"
