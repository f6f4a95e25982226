// oracle/orc_model.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Scalar restatement of how ybubnov/metalchat composes its kernels into the
// Llama-3 decode path: nn::linear / quantization::lora_linear / quantization::
// linear, nn::rmsnorm, nn::rope, nn::sink_cache, nn::attention, nn::feed_forward,
// nn::transformer, nn::llama3 and the top-k / nucleus / multinomial samplers.
// Every intermediate that the reference materialises in a T-typed buffer is
// rounded to T here at the same point (SURVEY.md §8a "rounding chain").
//
// PARITY STATUS: the op-level functions in orc_ops.h are pinned against the
// reference's own known-answer tests (tests/test_oracle_kat.py).  The reference
// has no fixtures for the rope rotation, attention, the transformer block, model
// logits or generated tokens (its integration tests only print text), so for the
// composition in this file the oracle is "parity unpinned": it is derived line by
// line from the cited headers and cross-checked against an independent fp32
// PyTorch Llama forward (tests/test_oracle_model.py).
#pragma once
#include "orc_ops.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace orc {

struct llama_cfg {
    uint32_t dim, n_layers, n_heads, n_kv_heads, head_dim, ffn_dim, vocab, max_seq_len;
    float rope_theta, norm_eps;
    uint32_t quant;      // 0: bf16/T weights, 1: QLoRA layout (huggingface/llama.h:152-171)
    uint32_t lora_rank;  // 16 (test/test_quantization.cc:50-51)
    float lora_scale;    // 2.0 (huggingface/llama.h:166-168)
    uint32_t group_size; // 32
    uint32_t n_seqs;     // independent bs=1 sequences (quirk Q2/Q15: max_batch_size = 1)
    uint32_t flags;      // ORC_* below
};
enum : uint32_t {
    ORC_UNTIED_HEAD = 1u << 0,   // T-typed model with its own `output` matrix (the reference always registers one, nn/llama.h:79;
                                 // huggingface/llama.h:103 aliases it to tok_embeddings when the checkpoint has no lm_head)
    ORC_PREFIX_VISIBLE = 1u << 1 // "intended" chunk mask: the cached prefix columns are 0 instead of -inf when len > 1 and
                                 // start_pos > 0 (the reference leaves them at -inf, quirk Q9, nn/attention.h:283-299)
};

// generator kinds (DESIGN.md "Synthetic data")
enum : uint32_t {
    K_ATTN_NORM = 0, K_FFN_NORM = 1, K_WQ = 2, K_WK = 3, K_WV = 4, K_WO = 5, K_W1 = 6, K_W2 = 7, K_W3 = 8,
    K_SCALES = 16, K_LORA_A = 32, K_LORA_B = 48,
    G_TOK = 0, G_NORM = 1, G_OUT = 2,
};
static inline uint64_t tid_layer(uint32_t layer, uint32_t kind) { return uint64_t(layer + 1) * 256 + kind; }
static inline uint64_t tid_global(uint32_t kind) { return kind; }

template <typename T> static inline T from_bf16_bits(uint16_t b) { return T(bf16_to_f32(b)); }

// One linear layer in any of the three weight formats.
template <typename T> struct linear_t {
    uint32_t N = 0, K = 0;
    int mode = 0; // 0: T weight; 1: lora_linear (int8 + group scales + LoRA); 2: int8 + per-row scale
    std::vector<T> w;
    std::vector<int8_t> q;
    std::vector<float> scales;
    std::vector<T> A, B;
    uint32_t rank = 0, group = 0;
    T lora_scale = T(0.0f);

    // dequantised row n as fp32 values of T: T(T(q) * T(s))  (kernel/mul.metal:76-77)
    void row_f32(uint32_t n, float* dst) const
    {
        if (mode == 0) {
            const T* r = &w[size_t(n) * K];
            for (uint32_t k = 0; k < K; k++) dst[k] = float(r[k]);
        } else if (mode == 1) {
            const int8_t* r = &q[size_t(n) * K];
            const float* s = &scales[size_t(n) * (K / group)];
            for (uint32_t k = 0; k < K; k++) {
                dst[k] = float(T(float(T(float(r[k]))) * float(T(s[k / group]))));
            }
        } else {
            const int8_t* r = &q[size_t(n) * K];
            const T s = T(scales[n]);
            for (uint32_t k = 0; k < K; k++) dst[k] = float(T(float(T(float(r[k]))) * float(s)));
        }
    }
};

// y[M,N] = T( sum_k x[m,k] * W[n,k] ), fp32 ascending-k accumulation from 0.0f
// (kernel/bmm.metal:55-76 through nn/linear.h:70-81).  Rows are processed eight at a
// time only to hide the add latency; each row keeps its own ascending-k chain.
template <typename T>
static void matmul_nt(const linear_t<T>& L, const float* wrows_or_null, const T* x, uint32_t M, T* y)
{
    (void)wrows_or_null;
    const uint32_t N = L.N, K = L.K;
    std::vector<float> xf(size_t(M) * K);
    for (size_t i = 0; i < xf.size(); i++) xf[i] = float(x[i]);
#pragma omp parallel
    {
        std::vector<float> wr(size_t(8) * K);
#pragma omp for schedule(static)
        for (uint32_t n0 = 0; n0 < N; n0 += 8) {
            const uint32_t nb = std::min<uint32_t>(8, N - n0);
            for (uint32_t r = 0; r < 8; r++) {
                if (r < nb) {
                    L.row_f32(n0 + r, &wr[size_t(r) * K]);
                } else {
                    std::fill(&wr[size_t(r) * K], &wr[size_t(r + 1) * K], 0.0f);
                }
            }
            const float *w0 = &wr[0], *w1 = &wr[K], *w2 = &wr[2 * size_t(K)], *w3 = &wr[3 * size_t(K)];
            const float *w4 = &wr[4 * size_t(K)], *w5 = &wr[5 * size_t(K)], *w6 = &wr[6 * size_t(K)],
                        *w7 = &wr[7 * size_t(K)];
            for (uint32_t m = 0; m < M; m++) {
                const float* xv = &xf[size_t(m) * K];
                float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
                for (uint32_t k = 0; k < K; k++) {
                    const float v = xv[k];
                    a0 += v * w0[k];
                    a1 += v * w1[k];
                    a2 += v * w2[k];
                    a3 += v * w3[k];
                    a4 += v * w4[k];
                    a5 += v * w5[k];
                    a6 += v * w6[k];
                    a7 += v * w7[k];
                }
                const float acc[8] = {a0, a1, a2, a3, a4, a5, a6, a7};
                for (uint32_t r = 0; r < nb; r++) y[size_t(m) * N + n0 + r] = T(acc[r]);
            }
        }
    }
}

// nn::linear (nn/linear.h:70-81), quantization::lora_linear (quantization/lora.h:
// 94-122) and quantization::linear (quantization/linear.h:50-57).
template <typename T> static void linear_forward(const linear_t<T>& L, const T* x, uint32_t M, T* y)
{
    matmul_nt(L, nullptr, x, M, y);
    if (L.mode == 1) {
        // adaptation = scalar_mul(B(A(x)), scale); result = add(output, adaptation)
        linear_t<T> la, lb;
        la.N = L.rank, la.K = L.K, la.w = L.A;
        lb.N = L.N, lb.K = L.rank, lb.w = L.B;
        std::vector<T> ax(size_t(M) * L.rank), bx(size_t(M) * L.N);
        matmul_nt(la, nullptr, x, M, ax.data());
        matmul_nt(lb, nullptr, ax.data(), M, bx.data());
        for (size_t i = 0; i < bx.size(); i++) {
            const T ad = T(float(bx[i]) * float(L.lora_scale));
            y[i] = T(float(y[i]) + float(ad));
        }
    }
}

// nn::attention::operator() from the scores on (nn/attention.h:195-203): for every (position t, head h)
//   s = T(q . K[s]) -> T(s * scale) -> [T(s + mask)] -> softmax without max shift (kernel/softmax.metal:40-80) -> o = T(sum_s p[s] V[s]).
// qr [len, H, hd] rotated queries; Kc / Vc [S.., KV, hd] cache rows 0..S-1; mask [len, S] or null; o [len, H, hd].
// repeat_interleave (functional/transform.h:80-91) => kv head = h / (H / KV).
template <typename T>
static void sdpa(const T* qr, const T* Kc, const T* Vc, uint32_t len, uint32_t S, uint32_t H, uint32_t KV, uint32_t hd, const T* mask, T scale, T* o)
{
    const uint32_t reps = H / KV;
    const uint32_t sm_block = ceil_div(S, kMaxThreads);
#pragma omp parallel for collapse(2) schedule(static)
    for (uint32_t hh = 0; hh < H; hh++) {
        for (uint32_t t = 0; t < len; t++) {
            const uint32_t kvh = hh / reps;
            std::vector<T> sc(S), pr(S);
            const T* qv = &qr[(size_t(t) * H + hh) * hd];
            for (uint32_t s = 0; s < S; s++) {
                const T* kv = &Kc[(size_t(s) * KV + kvh) * hd];
                float acc = 0.0f;
                for (uint32_t dd = 0; dd < hd; dd++) acc += float(qv[dd]) * float(kv[dd]);
                T sv = T(acc);                             // bmm
                sv = T(float(sv) * float(scale));          // scalar_mul
                if (mask) sv = T(float(sv) + float(mask[size_t(t) * S + s])); // add_broadcast
                sc[s] = sv;
            }
            layout<2> ls{{1, S}, {S, 1}, {0, 0}};
            softmax(pr.data(), ls, sc.data(), ls, sm_block);
            T* ov = &o[(size_t(t) * H + hh) * hd];
            for (uint32_t dd = 0; dd < hd; dd++) {
                float acc = 0.0f;
                for (uint32_t s = 0; s < S; s++) acc += float(pr[s]) * float(Vc[(size_t(s) * KV + kvh) * hd + dd]);
                ov[dd] = T(acc);
            }
        }
    }
}
// make_causal_mask (nn/attention.h:283-299): [len, S] filled with -inf, the trailing len x len square gets triu(diagonal = 1), i.e.
// 0 on and below the diagonal; the columns of the cached prefix stay at -inf (quirk Q9) unless `prefix_visible`.
template <typename T> static std::vector<T> causal_mask(uint32_t len, uint32_t S, bool prefix_visible)
{
    const T ninf = T(-std::numeric_limits<float>::infinity());
    std::vector<T> mask(size_t(len) * S, ninf);
    for (uint32_t i = 0; i < len; i++) {
        for (uint32_t j = 0; j <= i; j++) mask[size_t(i) * S + (S - len) + j] = T(0.0f);
        if (prefix_visible)
            for (uint32_t j = 0; j < S - len; j++) mask[size_t(i) * S + j] = T(0.0f);
    }
    return mask;
}

template <typename T> struct layer_t {
    std::vector<T> attn_norm, ffn_norm;
    linear_t<T> wq, wk, wv, wo, w1, w2, w3;
};

template <typename T> struct llama {
    llama_cfg cfg;
    std::vector<layer_t<T>> layers;
    linear_t<T> tok; // embedding table in linear_t form (mode 0 or 2)
    linear_t<T> out; // output head; mode 0 aliases tok (huggingface/llama.h:103)
    bool tied = true;
    std::vector<T> norm;
    std::vector<float> fcos, fsin;        // [2*max_seq, hd/2]  (nn/embedding.h:171), rows = positions rope_start + i
    uint32_t rope_start = 0;              // nn::rope::_M_start_pos (nn/embedding.h:160-199)
    std::vector<std::vector<T>> kc, vc;   // [seq*layer] -> [max_seq, n_kv, hd]
    std::vector<T> last_hidden;           // hidden state after the last block, [len, dim]
    std::unordered_map<std::string, std::pair<void*, size_t>> named;

    explicit llama(const llama_cfg& c) : cfg(c)
    {
        if (cfg.n_seqs == 0) cfg.n_seqs = 1;
        const uint32_t D = cfg.dim, H = cfg.n_heads, KV = cfg.n_kv_heads, hd = cfg.head_dim, F = cfg.ffn_dim;
        auto setup = [&](linear_t<T>& l, uint32_t N, uint32_t K, int mode) {
            l.N = N, l.K = K, l.mode = mode;
            if (mode == 0) {
                l.w.assign(size_t(N) * K, T(0.0f));
            } else if (mode == 1) {
                l.rank = cfg.lora_rank, l.group = cfg.group_size, l.lora_scale = T(cfg.lora_scale);
                l.q.assign(size_t(N) * K, 0);
                l.scales.assign(size_t(N) * (K / l.group), 0.0f);
                l.A.assign(size_t(l.rank) * K, T(0.0f));
                l.B.assign(size_t(N) * l.rank, T(0.0f));
            } else {
                l.q.assign(size_t(N) * K, 0);
                l.scales.assign(N, 0.0f);
            }
        };
        const int lm = cfg.quant ? 1 : 0;
        layers.resize(cfg.n_layers);
        for (auto& l : layers) {
            l.attn_norm.assign(D, T(0.0f));
            l.ffn_norm.assign(D, T(0.0f));
            setup(l.wq, H * hd, D, lm);
            setup(l.wk, KV * hd, D, lm);
            setup(l.wv, KV * hd, D, lm);
            setup(l.wo, D, H * hd, lm);
            setup(l.w1, F, D, lm);
            setup(l.w2, D, F, lm);
            setup(l.w3, F, D, lm);
        }
        setup(tok, cfg.vocab, D, cfg.quant ? 2 : 0);
        tied = !cfg.quant && !(cfg.flags & ORC_UNTIED_HEAD);
        if (!tied) setup(out, cfg.vocab, D, cfg.quant ? 2 : 0);
        norm.assign(D, T(0.0f));
        const uint32_t rows = 2 * cfg.max_seq_len;
        fcos.resize(size_t(rows) * (hd / 2));
        fsin.resize(size_t(rows) * (hd / 2));
        layout<2> lf{{rows, hd / 2}, {hd / 2, 1}, {0, 0}};
        rope_freqs(fcos.data(), lf, fsin.data(), lf, hd, 0, cfg.rope_theta); // nn/embedding.h:160-176
        kc.resize(size_t(cfg.n_seqs) * cfg.n_layers);
        vc.resize(size_t(cfg.n_seqs) * cfg.n_layers);
        for (auto& c2 : kc) c2.assign(size_t(cfg.max_seq_len) * KV * hd, T(0.0f));
        for (auto& c2 : vc) c2.assign(size_t(cfg.max_seq_len) * KV * hd, T(0.0f));
        register_names();
    }

    const linear_t<T>& head() const { return tied ? tok : out; }

    template <typename V> void reg(const std::string& n, std::vector<V>& v)
    {
        if (!v.empty()) named[n] = {static_cast<void*>(v.data()), v.size() * sizeof(V)};
    }
    void reg_linear(const std::string& p, linear_t<T>& l)
    {
        reg(p + ".weight", l.w);
        reg(p + ".weight", l.q);
        reg(p + ".scales", l.scales);
        reg(p + ".adaptor.A.weight", l.A);
        reg(p + ".adaptor.B.weight", l.B);
    }
    // Parameter names follow the reference's registered layer paths.
    void register_names()
    {
        for (uint32_t i = 0; i < cfg.n_layers; i++) {
            const std::string p = "layers." + std::to_string(i) + ".";
            reg(p + "attention_norm.weight", layers[i].attn_norm);
            reg(p + "ffn_norm.weight", layers[i].ffn_norm);
            reg_linear(p + "attention.wq", layers[i].wq);
            reg_linear(p + "attention.wk", layers[i].wk);
            reg_linear(p + "attention.wv", layers[i].wv);
            reg_linear(p + "attention.wo", layers[i].wo);
            reg_linear(p + "feed_forward.w1", layers[i].w1);
            reg_linear(p + "feed_forward.w2", layers[i].w2);
            reg_linear(p + "feed_forward.w3", layers[i].w3);
        }
        reg_linear("tok_embeddings", tok);
        if (!tied) reg_linear("output", out);
        reg("norm.weight", norm);
    }

    // ---- synthetic weights ------------------------------------------------------
    static void gen_T(std::vector<T>& v, uint64_t seed, uint64_t tid, float scale, float bias)
    {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < v.size(); i++) {
            const float u = hash_uniform(seed, tid, i);
            const float val = bias == 0.0f ? u * scale : bias + scale * u;
            v[i] = from_bf16_bits<T>(f32_to_bf16(val));
        }
    }
    static void gen_scales(std::vector<float>& v, uint64_t seed, uint64_t tid, float c)
    {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < v.size(); i++) {
            const float u = hash_uniform(seed, tid, i);
            v[i] = (1.0f + 0.5f * u) * c;
        }
    }
    static void gen_q(std::vector<int8_t>& v, uint64_t seed, uint64_t tid, int lo, uint32_t range)
    {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < v.size(); i++) v[i] = int8_t(hash_int(seed, tid, i, lo, range));
    }
    void gen_linear(linear_t<T>& l, uint64_t seed, uint64_t tid)
    {
        const float inv_sqrt_k = 1.0f / std::sqrt(float(l.K));
        if (l.mode == 0) {
            gen_T(l.w, seed, tid, inv_sqrt_k, 0.0f);
        } else {
            gen_q(l.q, seed, tid, -7, 15); // symmetric: a non-zero weight mean (UniformInt[-8,7] has -0.5) feeds a self-reinforcing all-ones mode through RMSNorm
                                           // and the hidden state degenerates to a constant vector whatever the tokens are
            gen_scales(l.scales, seed, tid + K_SCALES, inv_sqrt_k * 0.125f);
            gen_T(l.A, seed, tid + K_LORA_A, inv_sqrt_k, 0.0f);
            gen_T(l.B, seed, tid + K_LORA_B, 1.0f / std::sqrt(float(l.rank)), 0.0f);
        }
    }
    void init_random(uint64_t seed)
    {
        for (uint32_t i = 0; i < cfg.n_layers; i++) {
            auto& l = layers[i];
            gen_T(l.attn_norm, seed, tid_layer(i, K_ATTN_NORM), 0.1f, 1.0f);
            gen_T(l.ffn_norm, seed, tid_layer(i, K_FFN_NORM), 0.1f, 1.0f);
            gen_linear(l.wq, seed, tid_layer(i, K_WQ));
            gen_linear(l.wk, seed, tid_layer(i, K_WK));
            gen_linear(l.wv, seed, tid_layer(i, K_WV));
            gen_linear(l.wo, seed, tid_layer(i, K_WO));
            gen_linear(l.w1, seed, tid_layer(i, K_W1));
            gen_linear(l.w2, seed, tid_layer(i, K_W2));
            gen_linear(l.w3, seed, tid_layer(i, K_W3));
        }
        gen_T(norm, seed, tid_global(G_NORM), 0.1f, 1.0f);
        if (tok.mode == 0) {
            gen_T(tok.w, seed, tid_global(G_TOK), 0.0625f, 0.0f);
            if (!tied) gen_T(out.w, seed, tid_global(G_OUT), 1.0f / std::sqrt(float(cfg.dim)), 0.0f);
        } else {
            gen_q(tok.q, seed, tid_global(G_TOK), -127, 255);
            gen_scales(tok.scales, seed, tid_global(G_TOK) + K_SCALES, 0.0625f / 127.0f);
            gen_q(out.q, seed, tid_global(G_OUT), -127, 255);
            gen_scales(out.scales, seed, tid_global(G_OUT) + K_SCALES, (1.0f / std::sqrt(float(cfg.dim))) / 127.0f);
        }
    }

    // ---- forward ------------------------------------------------------------------
    void rmsnorm_rows(const T* x, uint32_t rows, const std::vector<T>& w, T* y) const
    {
        const uint32_t D = cfg.dim;
        layout<2> l2{{rows, D}, {D, 1}, {0, 0}};
        layout<1> l1{{D}, {1}, {0}};
        rmsnorm(y, l2, x, l2, w.data(), l1, cfg.norm_eps, 0.0f, ceil_div(D, kMaxThreads));
    }

    // nn::llama3::operator() (nn/llama.h:113-134) for one bs=1 sequence.
    // logits (T[vocab]) may be null.
    void forward(uint32_t seq, const int32_t* ids, uint32_t len, uint32_t start_pos, T* logits)
    {
        const uint32_t D = cfg.dim, H = cfg.n_heads, KV = cfg.n_kv_heads, hd = cfg.head_dim, F = cfg.ffn_dim;
        const uint32_t half = hd / 2;
        if (seq >= cfg.n_seqs) throw std::invalid_argument("oracle: sequence index out of range");
        // nn::sink_cache::copy (nn/cache.h:167-216): once start_pos has left the cache, the first pre_len = log2(max_seq_len) rows
        // (the "sink" tokens, nn/cache.h:123-126) stay, the other rows are rolled left by len (kernel/roll.metal:22-45) and the new
        // rows go to the end.  A call that straddles the end (start_pos < max_seq_len < start_pos + len) slices past the cache in the
        // reference; it is rejected here.
        const uint32_t cache_size = cfg.max_seq_len;
        const bool overflow = start_pos >= cache_size;
        if (len > cache_size) throw std::invalid_argument("sink_cache: requested length is larger than the cache size");
        if (!overflow && start_pos + len > cache_size) throw std::invalid_argument("oracle: the call straddles the end of the cache (undefined in the reference)");
        const uint32_t write_pos = overflow ? cache_size - len : start_pos;
        // embedding (nn/embedding.h:82-86; lora_embedding quantization/lora.h:160-170)
        std::vector<T> x(size_t(len) * D);
        {
            std::vector<float> row(D);
            for (uint32_t t = 0; t < len; t++) {
                const int32_t id = ids[t];
                if (id < 0 || uint32_t(id) >= cfg.vocab) throw std::invalid_argument("oracle: token id out of range");
                tok.row_f32(uint32_t(id), row.data());
                for (uint32_t k = 0; k < D; k++) x[size_t(t) * D + k] = T(row[k]);
            }
        }
        // make_causal_mask (nn/attention.h:283-299): only when len > 1; columns of the
        // cached prefix stay at -inf (quirk Q9).
        const uint32_t S = write_pos + len; // keys visible: cache[0, end_pos) (nn/cache.h:207-214); the mask is [len, min(start_pos + len, max_seq_len)] (nn/llama.h:119-121)
        std::vector<T> mask;
        if (len > 1) mask = causal_mask<T>(len, S, (cfg.flags & ORC_PREFIX_VISIBLE) != 0);
        const T scale = T(1.0f / std::sqrt(float(hd))); // stored as T (nn/attention.h:88,115; quirk Q4)

        std::vector<T> n(size_t(len) * D), q(size_t(len) * H * hd), k(size_t(len) * KV * hd), v(size_t(len) * KV * hd);
        std::vector<T> qr(q.size()), kr(k.size()), o(size_t(len) * H * hd), a(size_t(len) * D), h(size_t(len) * D);
        std::vector<T> g(size_t(len) * F), u(size_t(len) * F), z(size_t(len) * F), d(size_t(len) * D);

        for (uint32_t li = 0; li < cfg.n_layers; li++) {
            const layer_t<T>& L = layers[li];
            rmsnorm_rows(x.data(), len, L.attn_norm, n.data());
            linear_forward(L.wq, n.data(), len, q.data());
            linear_forward(L.wk, n.data(), len, k.data());
            linear_forward(L.wv, n.data(), len, v.data());
            // rope over [bs*len*n_head, hd] rows (kernel/embedding.h:87-125) at the ABSOLUTE position; the table covers 2 * max_seq_len
            // positions and is regenerated from start_pos when the position leaves it (nn/embedding.h:193-198)
            if (start_pos < rope_start || start_pos >= rope_start + 2 * cfg.max_seq_len) {
                rope_start = start_pos;
                layout<2> lt{{2 * cfg.max_seq_len, half}, {half, 1}, {0, 0}};
                rope_freqs(fcos.data(), lt, fsin.data(), lt, hd, rope_start, cfg.rope_theta);
            }
            {
                layout<2> lf{{2 * cfg.max_seq_len, half}, {half, 1}, {0, 0}};
                layout<2> lq{{len * H, hd}, {hd, 1}, {0, 0}};
                layout<2> lk{{len * KV, hd}, {hd, 1}, {0, 0}};
                rope(qr.data(), lq, q.data(), lq, fcos.data(), lf, fsin.data(), lf, 1, H, start_pos - rope_start);
                rope(kr.data(), lk, k.data(), lk, fcos.data(), lf, fsin.data(), lf, 1, KV, start_pos - rope_start);
            }
            // sink_cache::update (nn/cache.h:133-151,167-216): roll when full, then a bit copy into [write_pos, S)
            std::vector<T>& Kc = kc[size_t(seq) * cfg.n_layers + li];
            std::vector<T>& Vc = vc[size_t(seq) * cfg.n_layers + li];
            if (overflow) {
                uint32_t pre_len = 0; // std::bit_width(max_seq_len) - 1
                while ((2u << pre_len) <= cache_size) pre_len++;
                const uint32_t post_len = cache_size - pre_len;
                const size_t row = size_t(KV) * hd;
                for (std::vector<T>* c : {&Kc, &Vc}) {
                    std::vector<T> post(c->begin() + size_t(pre_len) * row, c->begin() + size_t(cache_size) * row);
                    for (uint32_t i = 0; i < post_len; i++) // out[i] = in[(i + shift) % size] along the position axis (kernel/roll.metal:36-41)
                        std::copy(post.begin() + size_t((i + len) % post_len) * row, post.begin() + size_t((i + len) % post_len + 1) * row,
                                  c->begin() + size_t(pre_len + i) * row);
                }
            }
            std::copy(kr.begin(), kr.end(), Kc.begin() + size_t(write_pos) * KV * hd);
            std::copy(v.begin(), v.end(), Vc.begin() + size_t(write_pos) * KV * hd);
            // attention (nn/attention.h:161-206)
            sdpa(qr.data(), Kc.data(), Vc.data(), len, S, H, KV, hd, len > 1 ? mask.data() : nullptr, scale, o.data());
            linear_forward(L.wo, o.data(), len, a.data());
            for (size_t i = 0; i < h.size(); i++) h[i] = T(float(x[i]) + float(a[i])); // nn/transformer.h:133
            rmsnorm_rows(h.data(), len, L.ffn_norm, n.data());
            linear_forward(L.w1, n.data(), len, g.data());
            linear_forward(L.w3, n.data(), len, u.data());
            for (size_t i = 0; i < z.size(); i++) z[i] = T(float(silu1(g[i])) * float(u[i])); // silu, hadamard
            linear_forward(L.w2, z.data(), len, d.data());
            for (size_t i = 0; i < x.size(); i++) x[i] = T(float(h[i]) + float(d[i])); // nn/transformer.h:139
        }
        last_hidden = x;
        if (logits) {
            // final norm over every position, then only the last one is projected
            // (nn/llama.h:128-133, quirk Q15)
            rmsnorm_rows(x.data(), len, norm, n.data());
            linear_forward(head(), &n[size_t(len - 1) * D], 1, logits);
        }
    }
};

// ---- samplers (nn/sampling.h) ------------------------------------------------------
struct sample_result {
    std::vector<int32_t> topk_idx;   // [k]   candidate ids after top-k (descending logit)
    std::vector<float> probs_sorted; // [k]   nucleus output (masked, descending), as fp32 values of T
    std::vector<int32_t> probs_idx;  // [k]   token ids aligned with probs_sorted
    int32_t choice = 0;              // index drawn by multinomial within [0,k)
    int32_t token = 0;               // sampled token id
};

// make_default_sampler (nn/sampling.h:306-316): top-k(k) -> nucleus(temperature, p) ->
// multinomial(1).  `u` is the injected uniform draw; `intended` selects the
// a = input[row, N-1] reading of the multinomial kernel (quirk Q10).  Top-k ties are
// broken by lower index first (the reference's heap order is unspecified, Q13).
template <typename T>
sample_result sample_default(const T* logits, uint32_t vocab, uint32_t topk, float temperature, float top_p, float u, int intended)
{
    sample_result r;
    const uint32_t k = std::min(vocab, topk);
    std::vector<int32_t> idx(vocab);
    for (uint32_t i = 0; i < vocab; i++) idx[i] = int32_t(i);
    std::partial_sort(idx.begin(), idx.begin() + k, idx.end(), [&](int32_t a, int32_t b) {
        const float fa = float(logits[a]), fb = float(logits[b]);
        return fa > fb || (fa == fb && a < b);
    });
    idx.resize(k);
    r.topk_idx = idx;
    std::vector<T> lg(k);
    for (uint32_t i = 0; i < k; i++) lg[i] = logits[idx[i]];
    // nucleus_sampler::sample (nn/sampling.h:183-200)
    const T temp_t = T(temperature);
    const T inv_t = T(1.0f / float(temp_t));
    const T p_t = T(top_p);
    layout<2> lk{{1, k}, {k, 1}, {0, 0}};
    std::vector<T> scaled(k), probs(k);
    scalar_mul(scaled.data(), lk, lg.data(), lk, inv_t);
    softmax(probs.data(), lk, scaled.data(), lk, ceil_div(k, kMaxThreads));
    const uint32_t P = ceil_pow2(k);
    layout<2> lp{{1, P}, {P, 1}, {0, 0}};
    std::vector<T> sv(P);
    std::vector<int32_t> si(P);
    sort(sv.data(), lp, si.data(), lp, probs.data(), lk);
    std::vector<T> cs(k), diff(k);
    const uint32_t cblock = std::max<uint32_t>(2, ceil_pow2(ceil_div(k, kMaxThreads)));
    cumsum(cs.data(), lk, sv.data(), lk, cblock);
    r.probs_sorted.resize(k);
    r.probs_idx.resize(k);
    std::vector<T> ps(k);
    for (uint32_t i = 0; i < k; i++) {
        diff[i] = T(float(cs[i]) - float(sv[i]));                // sub
        const bool m = float(diff[i]) > float(p_t);              // gt
        ps[i] = m ? T(0.0f) : sv[i];                             // scatter
        r.probs_sorted[i] = float(ps[i]);
        r.probs_idx[i] = idx[si[i]];                             // gather(context.indices, probs_idx)
    }
    // multinomial_sampler::sample (nn/sampling.h:289-297), sample_size = 1
    int32_t choice = 0;
    layout<2> lo{{1, 1}, {1, 1}, {0, 0}};
    multinomial(&choice, lo, ps.data(), lk, 0, 0, &u, intended);
    r.choice = choice;
    r.token = r.probs_idx[choice];
    return r;
}

// "greedy": argmax over logits, lowest index on ties (after Q17).
template <typename T> int32_t argmax(const T* logits, uint32_t vocab)
{
    uint32_t best = 0;
    float bv = float(logits[0]);
    for (uint32_t i = 1; i < vocab; i++) {
        const float f = float(logits[i]);
        if (f > bv) bv = f, best = i;
    }
    return int32_t(best);
}

} // namespace orc
