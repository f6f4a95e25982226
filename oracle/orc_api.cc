// oracle/orc_api.cc — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// C entry points of the CPU oracle for ctypes (tests/, bench.py's cpu_baseline and
// --impl reference legs, __graft_entry__.smoke()).  The product library
// (metalchat_b200/libmc_cuda.so) never links, loads or calls anything in here.
//
// dtype: 0 = "bfloat" (uint16 storage), 1 = "float".  Layout arguments are the
// reference's tensor_layout<N> PODs: 3*N uint32 {sizes, strides, offsets}.
#include "orc_model.h"

#include <chrono>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {
thread_local std::string g_err;
template <int N> const layout<N>& L(const uint32_t* p) { return *reinterpret_cast<const layout<N>*>(p); }
using bf = bf16_t;
} // namespace

#define ORC_DISPATCH(dtype, expr_bf, expr_f)  \
    do {                                      \
        if ((dtype) == 0) {                   \
            expr_bf;                          \
        } else {                              \
            expr_f;                           \
        }                                     \
    } while (0)

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_num_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

uint16_t orc_f32_to_bf16(float f) { return f32_to_bf16(f); }
uint16_t orc_f32_to_bf16_host(float f) { return f32_to_bf16_host(f); }
float orc_bf16_to_f32(uint16_t b) { return bf16_to_f32(b); }
uint64_t orc_hash3(uint64_t seed, uint64_t tid, uint64_t idx) { return hash3(seed, tid, idx); }
float orc_hash_uniform(uint64_t seed, uint64_t tid, uint64_t idx) { return hash_uniform(seed, tid, idx); }
int32_t orc_hash_int(uint64_t seed, uint64_t tid, uint64_t idx, int32_t lo, uint32_t range)
{
    return hash_int(seed, tid, idx, lo, range);
}

void orc_bmm(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, const void* b, const uint32_t* lb)
{
    ORC_DISPATCH(dt, bmm((bf*)o, L<3>(lo), (const bf*)a, L<3>(la), (const bf*)b, L<3>(lb)),
                 bmm((float*)o, L<3>(lo), (const float*)a, L<3>(la), (const float*)b, L<3>(lb)));
}
void orc_rmsnorm(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, const void* w, const uint32_t* lw, float eps, float mu, uint32_t block)
{
    ORC_DISPATCH(dt, rmsnorm((bf*)o, L<2>(lo), (const bf*)a, L<2>(la), (const bf*)w, L<1>(lw), eps, mu, block),
                 rmsnorm((float*)o, L<2>(lo), (const float*)a, L<2>(la), (const float*)w, L<1>(lw), eps, mu, block));
}
void orc_softmax(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, uint32_t block)
{
    ORC_DISPATCH(dt, softmax((bf*)o, L<2>(lo), (const bf*)a, L<2>(la), block),
                 softmax((float*)o, L<2>(lo), (const float*)a, L<2>(la), block));
}
void orc_sum(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, uint32_t block)
{
    ORC_DISPATCH(dt, sum((bf*)o, L<1>(lo), (const bf*)a, L<2>(la), block),
                 sum((float*)o, L<1>(lo), (const float*)a, L<2>(la), block));
}
void orc_rope(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, const float* c, const uint32_t* lc, const float* s, const uint32_t* ls, uint32_t bs, uint32_t n_head, uint32_t start_pos)
{
    ORC_DISPATCH(dt, rope((bf*)o, L<2>(lo), (const bf*)a, L<2>(la), c, L<2>(lc), s, L<2>(ls), bs, n_head, start_pos),
                 rope((float*)o, L<2>(lo), (const float*)a, L<2>(la), c, L<2>(lc), s, L<2>(ls), bs, n_head, start_pos));
}
void orc_rope_freqs(float* c, const uint32_t* lc, float* s, const uint32_t* ls, uint32_t dim, uint32_t start_pos, float theta)
{
    rope_freqs(c, L<2>(lc), s, L<2>(ls), dim, start_pos, theta);
}
void orc_embedding(int dt, void* o, const uint32_t* lo, const int32_t* ids, const uint32_t* li, const void* w, const uint32_t* lw)
{
    ORC_DISPATCH(dt, embedding((bf*)o, L<3>(lo), ids, L<2>(li), (const bf*)w, L<2>(lw)),
                 embedding((float*)o, L<3>(lo), ids, L<2>(li), (const float*)w, L<2>(lw)));
}
void orc_sort(int dt, void* v, const uint32_t* lv, int32_t* ix, const uint32_t* lx, const void* a, const uint32_t* la)
{
    ORC_DISPATCH(dt, sort((bf*)v, L<2>(lv), ix, L<2>(lx), (const bf*)a, L<2>(la)),
                 sort((float*)v, L<2>(lv), ix, L<2>(lx), (const float*)a, L<2>(la)));
}
void orc_cumsum(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, uint32_t block)
{
    ORC_DISPATCH(dt, cumsum((bf*)o, L<2>(lo), (const bf*)a, L<2>(la), block),
                 cumsum((float*)o, L<2>(lo), (const float*)a, L<2>(la), block));
}
void orc_multinomial(int dt, int32_t* o, const uint32_t* lo, const void* a, const uint32_t* la, uint64_t init_state, uint64_t init_seq, const float* uniforms, int intended)
{
    ORC_DISPATCH(dt, multinomial(o, L<2>(lo), (const bf*)a, L<2>(la), init_state, init_seq, uniforms, intended),
                 multinomial(o, L<2>(lo), (const float*)a, L<2>(la), init_state, init_seq, uniforms, intended));
}
float orc_pcg32_uniform(uint64_t init_state, uint64_t init_seq)
{
    pcg32 g(init_state, init_seq);
    return g.uniform();
}
// op: 0 add, 1 sub, 2 div, 3 hadamard
void orc_binary(int dt, int op, void* o, const uint32_t* lo, const void* a, const uint32_t* la, const void* b, const uint32_t* lb)
{
    auto run = [&](auto* po, const auto* pa, const auto* pb) {
        switch (op) {
        case 0: binary2(po, L<2>(lo), pa, L<2>(la), pb, L<2>(lb), [](float x, float y) { return x + y; }); break;
        case 1: binary2(po, L<2>(lo), pa, L<2>(la), pb, L<2>(lb), [](float x, float y) { return x - y; }); break;
        case 2: binary2(po, L<2>(lo), pa, L<2>(la), pb, L<2>(lb), [](float x, float y) { return x / y; }); break;
        default: binary2(po, L<2>(lo), pa, L<2>(la), pb, L<2>(lb), [](float x, float y) { return x * y; }); break;
        }
    };
    ORC_DISPATCH(dt, run((bf*)o, (const bf*)a, (const bf*)b), run((float*)o, (const float*)a, (const float*)b));
}
void orc_add_broadcast(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, const void* b, const uint32_t* lb)
{
    ORC_DISPATCH(dt, add_broadcast((bf*)o, L<2>(lo), (const bf*)a, L<2>(la), (const bf*)b, L<1>(lb)),
                 add_broadcast((float*)o, L<2>(lo), (const float*)a, L<2>(la), (const float*)b, L<1>(lb)));
}
// out dtype `odt`, scale dtype `sdt`
void orc_hadamard_broadcast(int odt, int sdt, void* o, const uint32_t* lo, const int8_t* a, const uint32_t* la, const void* b, const uint32_t* lb)
{
    if (odt == 0 && sdt == 0) hadamard_broadcast((bf*)o, L<2>(lo), a, L<2>(la), (const bf*)b, L<1>(lb));
    if (odt == 0 && sdt == 1) hadamard_broadcast((bf*)o, L<2>(lo), a, L<2>(la), (const float*)b, L<1>(lb));
    if (odt == 1 && sdt == 0) hadamard_broadcast((float*)o, L<2>(lo), a, L<2>(la), (const bf*)b, L<1>(lb));
    if (odt == 1 && sdt == 1) hadamard_broadcast((float*)o, L<2>(lo), a, L<2>(la), (const float*)b, L<1>(lb));
}
void orc_scalar_mul(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, float c)
{
    // `c` arrives as the fp32 value of the T-typed scalar the host bound.
    ORC_DISPATCH(dt, scalar_mul((bf*)o, L<2>(lo), (const bf*)a, L<2>(la), bf(c)),
                 scalar_mul((float*)o, L<2>(lo), (const float*)a, L<2>(la), c));
}
// op: 0 silu, 1 gelu
void orc_activation(int dt, int op, void* o, const uint32_t* lo, const void* a, const uint32_t* la)
{
    if (op == 0) {
        ORC_DISPATCH(dt, silu((bf*)o, L<2>(lo), (const bf*)a, L<2>(la)), silu((float*)o, L<2>(lo), (const float*)a, L<2>(la)));
    } else {
        ORC_DISPATCH(dt, gelu((bf*)o, L<2>(lo), (const bf*)a, L<2>(la)), gelu((float*)o, L<2>(lo), (const float*)a, L<2>(la)));
    }
}
// dt: 0 bf16, 1 f32, 2 int32
void orc_copy(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la)
{
    if (dt == 0) copy((uint16_t*)o, L<2>(lo), (const uint16_t*)a, L<2>(la));
    else copy((uint32_t*)o, L<2>(lo), (const uint32_t*)a, L<2>(la));
}
void orc_gather(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, const int32_t* ix, const uint32_t* li)
{
    if (dt == 0) gather((uint16_t*)o, L<2>(lo), (const uint16_t*)a, L<2>(la), ix, L<2>(li));
    else gather((uint32_t*)o, L<2>(lo), (const uint32_t*)a, L<2>(la), ix, L<2>(li));
}
void orc_scatter(int dt, void* o, const uint32_t* lo, const uint8_t* m, const uint32_t* lm, float value)
{
    ORC_DISPATCH(dt, scatter((bf*)o, L<2>(lo), m, L<2>(lm), bf(value)), scatter((float*)o, L<2>(lo), m, L<2>(lm), value));
}
// op: 0 gt, 1 le
void orc_compare(int dt, int op, uint8_t* o, const uint32_t* lo, const void* a, const uint32_t* la, float value)
{
    if (op == 0) {
        ORC_DISPATCH(dt, gt(o, L<2>(lo), (const bf*)a, L<2>(la), bf(value)), gt(o, L<2>(lo), (const float*)a, L<2>(la), value));
    } else {
        ORC_DISPATCH(dt, le(o, L<2>(lo), (const bf*)a, L<2>(la), bf(value)), le(o, L<2>(lo), (const float*)a, L<2>(la), value));
    }
}
void orc_roll(int dt, void* o, const uint32_t* lo, const void* a, const uint32_t* la, uint32_t shift, uint32_t size, uint32_t stride)
{
    if (dt == 0) roll((uint16_t*)o, L<1>(lo), (const uint16_t*)a, L<1>(la), shift, size, stride);
    else roll((uint32_t*)o, L<1>(lo), (const uint32_t*)a, L<1>(la), shift, size, stride);
}

// torchrun exports OMP_NUM_THREADS=1; the CPU-baseline legs of bench.py size the pool explicitly
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// ---- model ------------------------------------------------------------------------
struct orc_model {
    int dtype;
    std::unique_ptr<llama<bf>> b;
    std::unique_ptr<llama<float>> f;
};

void* orc_llama_create(const llama_cfg* cfg, int dtype)
{
    try {
        auto* m = new orc_model();
        m->dtype = dtype;
        if (dtype == 0) m->b = std::make_unique<llama<bf>>(*cfg);
        else m->f = std::make_unique<llama<float>>(*cfg);
        return m;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void orc_llama_destroy(void* h) { delete static_cast<orc_model*>(h); }
void orc_llama_init_random(void* h, uint64_t seed)
{
    auto* m = static_cast<orc_model*>(h);
    if (m->b) m->b->init_random(seed);
    else m->f->init_random(seed);
}
// Raw storage of a named parameter (reference layer paths); nbytes receives its size.
void* orc_llama_tensor(void* h, const char* name, uint64_t* nbytes)
{
    auto* m = static_cast<orc_model*>(h);
    auto& named = m->b ? m->b->named : m->f->named;
    auto it = named.find(name);
    if (it == named.end()) {
        g_err = std::string("oracle: no parameter named ") + name;
        return nullptr;
    }
    if (nbytes) *nbytes = it->second.second;
    return it->second.first;
}
// KV cache storage of (seq, layer): which = 0 keys, 1 values.
void* orc_llama_cache(void* h, uint32_t seq, uint32_t layer, int which, uint64_t* nbytes)
{
    auto* m = static_cast<orc_model*>(h);
    if (m->b) {
        auto& c = (which ? m->b->vc : m->b->kc)[size_t(seq) * m->b->cfg.n_layers + layer];
        if (nbytes) *nbytes = c.size() * 2;
        return c.data();
    }
    auto& c = (which ? m->f->vc : m->f->kc)[size_t(seq) * m->f->cfg.n_layers + layer];
    if (nbytes) *nbytes = c.size() * 4;
    return c.data();
}
// rc 0 ok. logits/hidden may be null. hidden receives [len, dim] after the last block.
int orc_llama_forward(void* h, uint32_t seq, const int32_t* ids, uint32_t len, uint32_t start_pos, void* logits, void* hidden)
{
    auto* m = static_cast<orc_model*>(h);
    try {
        if (m->b) {
            m->b->forward(seq, ids, len, start_pos, (bf*)logits);
            if (hidden) std::memcpy(hidden, m->b->last_hidden.data(), m->b->last_hidden.size() * 2);
        } else {
            m->f->forward(seq, ids, len, start_pos, (float*)logits);
            if (hidden) std::memcpy(hidden, m->f->last_hidden.data(), m->f->last_hidden.size() * 4);
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
// Stand-alone attention (scores -> scale -> mask -> softmax -> values) for the isolated parity tests of the attention kernels.
// q [len, H, hd] rotated; K / V [S, KV, hd]; causal != 0 builds the reference's mask for len > 1 (prefix_visible selects quirk Q9's
// intended reading); scale is stored as T(1 / sqrt(hd)) (nn/attention.h:88,115).
void orc_sdpa(int dt, void* o, const void* q, const void* K, const void* V, uint32_t len, uint32_t S, uint32_t H, uint32_t KV, uint32_t hd, int causal, int prefix_visible)
{
    if (dt == 0) {
        std::vector<bf> mask;
        if (causal && len > 1) mask = causal_mask<bf>(len, S, prefix_visible != 0);
        sdpa((const bf*)q, (const bf*)K, (const bf*)V, len, S, H, KV, hd, mask.empty() ? nullptr : mask.data(), bf(1.0f / std::sqrt(float(hd))), (bf*)o);
    } else {
        std::vector<float> mask;
        if (causal && len > 1) mask = causal_mask<float>(len, S, prefix_visible != 0);
        sdpa((const float*)q, (const float*)K, (const float*)V, len, S, H, KV, hd, mask.empty() ? nullptr : mask.data(), 1.0f / std::sqrt(float(hd)), (float*)o);
    }
}
int32_t orc_argmax(int dt, const void* logits, uint32_t vocab)
{
    return dt == 0 ? argmax((const bf*)logits, vocab) : argmax((const float*)logits, vocab);
}
// Default sampler chain; outputs: topk_idx[k], probs_sorted[k] (fp32 values of T),
// probs_idx[k], choice, token.  Returns k.
uint32_t orc_sample_default(int dt, const void* logits, uint32_t vocab, uint32_t topk, float temperature, float top_p, float u, int intended, int32_t* topk_idx, float* probs_sorted, int32_t* probs_idx, int32_t* choice, int32_t* token)
{
    sample_result r = dt == 0 ? sample_default((const bf*)logits, vocab, topk, temperature, top_p, u, intended)
                              : sample_default((const float*)logits, vocab, topk, temperature, top_p, u, intended);
    const uint32_t k = uint32_t(r.topk_idx.size());
    if (topk_idx) std::copy(r.topk_idx.begin(), r.topk_idx.end(), topk_idx);
    if (probs_sorted) std::copy(r.probs_sorted.begin(), r.probs_sorted.end(), probs_sorted);
    if (probs_idx) std::copy(r.probs_idx.begin(), r.probs_idx.end(), probs_idx);
    if (choice) *choice = r.choice;
    if (token) *token = r.token;
    return k;
}

// Timed greedy decode loop for the CPU baseline: runs `steps` decode steps starting
// at `start_pos` from token `first_id`, returns seconds; tokens (nullable) receives ids.
double orc_llama_decode_timed(void* h, int32_t first_id, uint32_t start_pos, uint32_t steps, int32_t* tokens)
{
    auto* m = static_cast<orc_model*>(h);
    const uint32_t vocab = m->b ? m->b->cfg.vocab : m->f->cfg.vocab;
    std::vector<uint8_t> logits(size_t(vocab) * 4);
    int32_t id = first_id;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t s = 0; s < steps; s++) {
        if (orc_llama_forward(h, 0, &id, 1, start_pos + s, logits.data(), nullptr)) return -1.0;
        id = orc_argmax(m->dtype, logits.data(), vocab);
        if (tokens) tokens[s] = id;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // extern "C"
