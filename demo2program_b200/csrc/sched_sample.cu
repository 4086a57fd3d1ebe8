// Scheduled sampling of the teacher-forced token decoders (reference models/model_full.py:59-67,
// 414-423: seq2seq.ScheduledEmbeddingTrainingHelper(embedding, seq_lengths, embedding_lookup,
// 1 - sample_prob) with sample_prob = polynomial_decay(1.0 -> 0.1 over
// scheduled_sampling_decay_steps, power 1) of the global step; enabled by trainer.py:278-281).
//
// After decoder step t every row draws, with probability p = 1 - sample_prob, the token it feeds to
// step t+1 from Categorical(logits_t) instead of taking the ground-truth token.  The reference leaves
// both draws unseeded; here they are a counter-based hash of (seed, global step, decoder, t, row),
// so a step is reproducible and the CPU oracle can restate it.  The sampled token is not
// differentiated through (it is an integer), so the backward pass is the teacher-forced one over the
// tokens that were actually fed.
#include "common.cuh"

namespace d2p {
namespace {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// which: 0 = the Bernoulli draw (sample or ground truth), 1 = the categorical draw
__host__ __device__ __forceinline__ uint32_t sched_hash(uint32_t seed, uint32_t step, uint32_t decoder, uint32_t t,
                                                        uint32_t r, uint32_t which) {
    uint32_t h = mix32(seed + 0x9E3779B9U);
    h = mix32(h ^ step);
    h = mix32(h + decoder);
    h = mix32(h ^ t);
    h = mix32(h + r);
    h = mix32(h ^ which);
    return h;
}

__global__ void step_lens_kernel(const int* __restrict__ runlen, int R, int L, int* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * L) return;
    const int t = idx / R, r = idx - t * R;
    out[idx] = t < runlen[r] ? 1 : 0;
}

// out[r, :] = table[id] with id = (t == 0 ? start_id : tokens[r, t-1]); out of range -> 0
__global__ void embed_step_kernel(const float4* __restrict__ table, int vocab_rows, int E4,
                                  const int* __restrict__ tokens, int R, int L, int t, int start_id,
                                  float4* __restrict__ out) {
    const size_t total = (size_t)R * E4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / E4), e = (int)(idx - (size_t)r * E4);
        const int id = t == 0 ? start_id : tokens[(size_t)r * L + t - 1];
        out[idx] = (id >= 0 && id < vocab_rows) ? table[(size_t)id * E4 + e] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void sched_sample_kernel(const float* __restrict__ logits, int R, int V, const int* __restrict__ gt,
                                    int L, int t, const double* __restrict__ adam_state, int decay_steps,
                                    float p_override, uint32_t seed, uint32_t decoder, int* __restrict__ fed,
                                    int* __restrict__ sampled) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const double gstep = adam_state ? adam_state[0] : 0.0;
    double p = p_override;
    if (p_override < 0.f) {
        // tf.train.polynomial_decay(1.0, global_step, decay_steps, end_learning_rate=0.1, power=1.0)
        const double frac = fmin(gstep, (double)decay_steps) / (double)decay_steps;
        const double teacher = (1.0 - 0.1) * (1.0 - frac) + 0.1;
        p = 1.0 - teacher;
    }
    const uint32_t step = (uint32_t)gstep;
    const double u1 = (double)sched_hash(seed, step, decoder, (uint32_t)t, (uint32_t)r, 0u) * (1.0 / 4294967296.0);
    const bool take = p > u1;                 // select_sample = sampling_probability > uniform
    int tok = gt[(size_t)r * L + t];
    if (take) {
        const float* x = logits + (size_t)r * V;
        float m = x[0];
        for (int v = 1; v < V; ++v) m = fmaxf(m, x[v]);
        float s = 0.f;
        for (int v = 0; v < V; ++v) s += expf(x[v] - m);
        const float u2 = (float)((double)sched_hash(seed, step, decoder, (uint32_t)t, (uint32_t)r, 1u) * (1.0 / 4294967296.0));
        const float target = u2 * s;
        float c = 0.f;
        tok = V - 1;
        for (int v = 0; v < V; ++v) {
            c += expf(x[v] - m);
            if (c > target) { tok = v; break; }
        }
    }
    fed[(size_t)r * L + t] = tok;
    if (sampled) sampled[(size_t)r * L + t] = take ? 1 : 0;
}

}  // namespace
}  // namespace d2p

using namespace d2p;

extern "C" int d2p_step_lens(const int* runlen, int R, int L, int* out, void* stream) {
    D2P_REQUIRE(runlen && out && R > 0 && L > 0, "step_lens: bad arguments");
    step_lens_kernel<<<cdiv((long long)R * L, 256), 256, 0, (cudaStream_t)stream>>>(runlen, R, L, out);
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_embed_shifted_step(const float* table, int vocab_rows, int E, const int* tokens, int R, int L, int t,
                                      int start_id, float* out, void* stream) {
    D2P_REQUIRE(table && tokens && out && R > 0 && t >= 0 && t < L, "embed_shifted_step: bad arguments");
    D2P_REQUIRE(E % 4 == 0 && (((uintptr_t)table | (uintptr_t)out) & 15) == 0,
                "embed_shifted_step: rows must be 16-byte aligned multiples of 4 floats");
    const size_t total = (size_t)R * (E / 4), nb = (total + 255) / 256, cap = 8 * (size_t)kNumSMs;
    embed_step_kernel<<<(int)(nb < cap ? nb : cap), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(table), vocab_rows, E / 4, tokens, R, L, t, start_id,
        reinterpret_cast<float4*>(out));
    D2P_CHECK_LAUNCH();
    return 0;
}

extern "C" int d2p_sched_sample_step(const float* logits, int R, int V, const int* gt_tokens, int L, int t,
                                     const double* adam_state, int decay_steps, float p_override, unsigned seed,
                                     int decoder, int* fed_tokens, int* sampled, void* stream) {
    D2P_REQUIRE(logits && gt_tokens && fed_tokens && R > 0 && V > 0 && t >= 0 && t < L,
                "sched_sample_step: bad arguments");
    D2P_REQUIRE(p_override >= 0.f || (adam_state && decay_steps > 0), "sched_sample_step: no schedule given");
    sched_sample_kernel<<<cdiv(R, 128), 128, 0, (cudaStream_t)stream>>>(logits, R, V, gt_tokens, L, t, adam_state,
                                                                         decay_steps, p_override, (uint32_t)seed,
                                                                         (uint32_t)decoder, fed_tokens, sampled);
    D2P_CHECK_LAUNCH();
    return 0;
}

/* host restatement of the draws (tests / oracle pin the hash against it without a GPU) */
extern "C" unsigned d2p_sched_hash(unsigned seed, unsigned step, unsigned decoder, unsigned t, unsigned r,
                                   unsigned which) {
    return sched_hash(seed, step, decoder, t, r, which);
}
