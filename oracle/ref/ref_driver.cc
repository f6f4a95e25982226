// oracle/ref/ref_driver.cc — TEST INFRASTRUCTURE ONLY.
//
// Runs the REFERENCE'S OWN layer code -- nn::llama3<bf16> (nn/llama.h:113-134), its QLoRA variant assembled by the reference's
// own huggingface::llama3_qlora_safetensor_serializer::adapt (huggingface/llama.h:152-171), nn::gemma3<bf16> (nn/gemma.h:43-147)
// and make_default_sampler (nn/sampling.h:306-316) -- on the synthetic hash weights of DESIGN.md "Synthetic data", through the
// façade (metalchat_b200/facade) and whichever backend of the C ABI this binary was linked with:
//
//   oracle/_ref/cpu/ref_driver   + liborc_mc.so  : the reference's composition over the oracle's op kernels on the host.
//                                                  tests/test_oracle_ref.py requires BIT-EQUALITY with oracle/orc_model.h --
//                                                  this pins the oracle's composition (attention, cache, rope, blocks, QLoRA
//                                                  rounding order, sampler chain) to the reference's own code.
//   oracle/_ref/cuda/ref_driver  + libmc_cuda.so : the same program on the B200 (tests/test_gpu_ref_driver.py): reference-style
//                                                  C++ user code running on the CUDA backend, compared with the CPU run.
//
// usage: ref_driver <llama|qlora|gemma> <small|hd128> <out.bin> <n_prompt> <n_decode> [chunk]
//   prompt ids = hash(0x5EED, 0xFFFF, i); `chunk` > 0 feeds the prompt in two calls (the second at start_pos = chunk: quirk Q9).
// output: "MCRF" | u32 vocab | u32 n_forward | per forward: vocab x u16 logits bits | n_forward x i32 token of the default sampler
//         | n_forward x i32 greedy argmax (lowest index on ties)
#include <cstdio>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include <metalchat/accelerator.h>
#include <metalchat/functional.h>
#include <metalchat/huggingface/llama.h>
#include <metalchat/nn.h>
#include <metalchat/nn/gemma.h>
#include <metalchat/nn/llama.h>
#include <metalchat/quantization.h>
#include <metalchat/tensor.h>

#include "orc_common.h"

using namespace metalchat;

namespace {

struct shape_cfg {
    std::size_t dim, n_layers, n_heads, n_kv_heads, head_dim, ffn, vocab, max_seq;
};
const shape_cfg kSmall{512, 3, 8, 2, 64, 1024, 2000, 96};
const shape_cfg kHd128{512, 2, 6, 2, 128, 768, 1500, 160};

// generator kinds of DESIGN.md "Synthetic data" (orc_model.h) + the extra norms of Gemma-3
enum : uint32_t { K_ATTN_NORM = 0, K_FFN_NORM = 1, K_WQ = 2, K_WK = 3, K_WV = 4, K_WO = 5, K_W1 = 6, K_W2 = 7, K_W3 = 8, K_Q_NORM = 9, K_K_NORM = 10,
                  K_ATTN_POST_NORM = 11, K_FFN_POST_NORM = 12, K_SCALES = 16, K_LORA_A = 32, K_LORA_B = 48, G_TOK = 0, G_NORM = 1, G_OUT = 2 };
constexpr uint64_t kSeed = 0x5EED;

struct spec {
    std::vector<std::size_t> sizes;
    uint64_t tid = 0;
    int gen = 0;          // 0: bf16 u * scale (+ bias), 1: int8 uniform, 2: fp32 group scales (1 + 0.5 u) * c
    float scale = 0.0f, bias = 0.0f;
    int lo = 0;
    uint32_t range = 0;
};

bool
ends_with(const std::string& s, const std::string& suffix)
{
    return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

// shape and generator of a registered parameter, from its dotted path
bool
parameter_spec(const std::string& path, const shape_cfg& c, bool quant, spec& out)
{
    const std::size_t D = c.dim, QO = c.n_heads * c.head_dim, KO = c.n_kv_heads * c.head_dim, F = c.ffn, V = c.vocab;
    auto linear = [&](const std::string& rest, std::size_t N, std::size_t K, uint64_t tid) {
        const float inv = 1.0f / std::sqrt(float(K));
        if (rest == "weight" && !quant) out = {{N, K}, tid, 0, inv, 0.0f};
        else if (rest == "weight") out = {{N, K}, tid, 1, 0, 0, -7, 15};
        else if (rest == "scales") out = {{N, K / 32}, tid + K_SCALES, 2, inv * 0.125f};
        else if (rest == "adaptor.A.weight") out = {{16, K}, tid + K_LORA_A, 0, inv, 0.0f};
        else if (rest == "adaptor.B.weight") out = {{N, 16}, tid + K_LORA_B, 0, 0.25f, 0.0f};
        else return false;
        return true;
    };
    if (path == "tok_embeddings.weight") {
        if (quant) out = {{V, D}, G_TOK, 1, 0, 0, -127, 255};
        else out = {{V, D}, G_TOK, 0, 0.0625f, 0.0f};
        return true;
    }
    if (path == "tok_embeddings.scales") return out = {{V, 1}, G_TOK + K_SCALES, 2, 0.0625f / 127.0f}, true;
    if (path == "output.weight") {
        if (quant) out = {{V, D}, G_OUT, 1, 0, 0, -127, 255};
        else out = {{V, D}, G_OUT, 0, 1.0f / std::sqrt(float(D)), 0.0f}; // only used by untied models (gemma driver)
        return true;
    }
    if (path == "output.scales") return out = {{V, 1}, G_OUT + K_SCALES, 2, (1.0f / std::sqrt(float(D))) / 127.0f}, true;
    if (path == "norm.weight") return out = {{D}, G_NORM, 0, 0.1f, 1.0f}, true;
    if (path.rfind("layers.", 0) != 0) return false;
    const auto dot = path.find('.', 7);
    const uint32_t li = uint32_t(std::stoul(path.substr(7, dot - 7)));
    const std::string rest = path.substr(dot + 1);
    auto tid = [&](uint32_t kind) { return uint64_t(li + 1) * 256 + kind; };
    if (rest == "attention_norm.weight") return out = {{D}, tid(K_ATTN_NORM), 0, 0.1f, 1.0f}, true;
    if (rest == "ffn_norm.weight") return out = {{D}, tid(K_FFN_NORM), 0, 0.1f, 1.0f}, true;
    // Gemma-3: weights are stored as (w - 1) because its RMSNorm adds mu = 1 (nn/gemma.h:79,106-107)
    if (rest == "attention_post_norm.weight") return out = {{D}, tid(K_ATTN_POST_NORM), 0, 0.1f, 0.0f}, true;
    if (rest == "ffn_post_norm.weight") return out = {{D}, tid(K_FFN_POST_NORM), 0, 0.1f, 0.0f}, true;
    if (rest == "attention.q_norm.weight") return out = {{c.head_dim}, tid(K_Q_NORM), 0, 0.1f, 0.0f}, true;
    if (rest == "attention.k_norm.weight") return out = {{c.head_dim}, tid(K_K_NORM), 0, 0.1f, 0.0f}, true;
    struct {
        const char* prefix;
        std::size_t N, K;
        uint32_t kind;
    } const lins[] = {{"attention.wq.", QO, D, K_WQ}, {"attention.wk.", KO, D, K_WK}, {"attention.wv.", KO, D, K_WV}, {"attention.wo.", D, QO, K_WO},
                      {"feed_forward.w1.", F, D, K_W1}, {"feed_forward.w2.", D, F, K_W2}, {"feed_forward.w3.", F, D, K_W3}};
    for (const auto& l : lins) {
        const std::string p = l.prefix;
        if (rest.rfind(p, 0) == 0) return linear(rest.substr(p.size()), l.N, l.K, tid(l.kind));
    }
    return false;
}

float
gen_uniform(uint64_t tid, uint64_t i)
{
    return orc::hash_uniform(kSeed, tid, i);
}

// Allocates hardware memory for a parameter, fills it with its synthetic values and re-points the registered tensor at it --
// what safetensor_document::load does with a file-backed container (src/safetensor.cc:237-253).
template <typename T, std::size_t N, typename Fill>
void
assign(basic_tensor& param, const std::vector<std::size_t>& sizes, hardware_accelerator& gpu, Fill fill)
{
    std::size_t dims[N];
    std::copy(sizes.begin(), sizes.end(), dims);
    auto t = empty<T>(std::move(dims), gpu.get_allocator());
    T* p = t.data_ptr();
    const std::size_t n = t.numel();
    for (std::size_t i = 0; i < n; i++) p[i] = fill(i);
    tensor_accessor::resize(t.sizes(), param);
    param.set_container(t.container_ptr());
}

template <typename Fill>
void
assign_any(basic_tensor& param, const spec& s, hardware_accelerator& gpu, const std::string& path, Fill dispatch)
{
    if (param.dimensions() != s.sizes.size()) throw std::runtime_error("parameter " + path + ": unexpected number of dimensions");
    dispatch(param, s, gpu);
}

void
fill_parameter(basic_tensor& param, const spec& s, hardware_accelerator& gpu, const std::string& path, float tok_mult)
{
    const auto& dt = param.dtype();
    auto bf = [&](std::size_t i) {
        const float u = gen_uniform(s.tid, i);
        float v = s.bias == 0.0f ? u * s.scale : s.bias + s.scale * u;
        bf16 r(orc::bf16_to_f32(orc::f32_to_bf16(v)));
        if (tok_mult != 1.0f) r = bf16(float(r) * tok_mult);
        return r;
    };
    auto i8 = [&](std::size_t i) { return std::int8_t(orc::hash_int(kSeed, s.tid, i, s.lo, s.range)); };
    auto f32 = [&](std::size_t i) { return (1.0f + 0.5f * gen_uniform(s.tid, i)) * s.scale; };
    if (param.dimensions() != s.sizes.size()) throw std::runtime_error("parameter " + path + ": unexpected number of dimensions");
    if (dt == typeid(bf16) && s.gen == 0) {
        if (s.sizes.size() == 1) assign<bf16, 1>(param, s.sizes, gpu, bf);
        else assign<bf16, 2>(param, s.sizes, gpu, bf);
    } else if (dt == typeid(std::int8_t) && s.gen == 1) {
        assign<std::int8_t, 2>(param, s.sizes, gpu, i8);
    } else if (dt == typeid(float) && s.gen == 2) {
        assign<float, 2>(param, s.sizes, gpu, f32);
    } else {
        throw std::runtime_error("parameter " + path + ": dtype does not match its generator");
    }
}

template <typename Layer>
void
init_parameters(Layer& model, const shape_cfg& c, bool quant, bool tie_output, hardware_accelerator& gpu)
{
    std::size_t n = 0;
    for (auto& named : model.parameters()) {
        const std::string& path = named.path;
        if (path.find(".cache.") != std::string::npos) continue; // runtime state registered as parameters (nn/cache.h:219-223)
        if (tie_output && path == "output.weight") continue;
        spec s;
        if (!parameter_spec(path, c, quant, s)) throw std::runtime_error("no synthetic generator for parameter " + path);
        fill_parameter(*named.ptr, s, gpu, path, 1.0f);
        n++;
    }
    if (tie_output) {
        // huggingface/llama.h:103: doc.insert("output.weight", "tok_embeddings.weight") -- the two parameters share one container
        auto& tok = model.parameter("tok_embeddings.weight");
        auto& out = model.parameter("output.weight");
        tensor_accessor::resize(tok.sizes(), out);
        out.set_container(tok.container_ptr());
    }
    std::fprintf(stderr, "ref_driver: %zu parameters initialised\n", n);
}

int32_t
argmax_lowest(const bf16* logits, std::size_t n)
{
    std::size_t best = 0;
    float bv = float(logits[0]);
    for (std::size_t i = 1; i < n; i++) {
        const float f = float(logits[i]);
        if (f > bv) bv = f, best = i;
    }
    return int32_t(best);
}

struct recorder {
    std::size_t vocab;
    std::vector<std::uint16_t> logits;
    std::vector<int32_t> sampled, greedy;
};

// one forward of `ids` at start_pos through the reference's layer, then the reference's default sampler on the logits
template <typename Model>
int32_t
step(Model& model, nn::basic_sampler<bf16>& sampler, hardware_accelerator& gpu, const std::vector<int32_t>& ids, std::size_t start_pos, recorder& rec, bool record)
{
    auto input = shared_tensor(to_tensor<int32_t>({1, ids.size()}, ids.begin(), ids.end()));
    auto logits = model(input, start_pos);            // [1, 1, vocab]
    auto flat = logits.template flatten<2>();         // transformer.h:357-364 samples from [1, vocab]
    auto token = sampler.sample(flat, gpu).get();     // nn/sampling.h:69-76
    auto host = logits.get();
    const bf16* p = host.data_ptr();
    const int32_t greedy = argmax_lowest(p, rec.vocab);
    if (record) {
        for (std::size_t i = 0; i < rec.vocab; i++) {
            std::uint16_t bits;
            std::memcpy(&bits, &p[i], 2);
            rec.logits.push_back(bits);
        }
        rec.sampled.push_back(token[0, 0]);
        rec.greedy.push_back(greedy);
    }
    return greedy;
}

template <typename Model>
void
run(Model& model, hardware_accelerator& gpu, const shape_cfg& c, std::size_t n_prompt, std::size_t n_decode, std::size_t chunk, recorder& rec)
{
    auto sampler = nn::make_default_sampler<bf16>();
    std::vector<int32_t> prompt(n_prompt);
    for (std::size_t i = 0; i < n_prompt; i++) prompt[i] = orc::hash_int(kSeed, 0xFFFF, i, 0, uint32_t(c.vocab));
    int32_t tok;
    if (chunk > 0 && chunk < n_prompt) {
        step(model, *sampler, gpu, std::vector<int32_t>(prompt.begin(), prompt.begin() + chunk), 0, rec, true);
        tok = step(model, *sampler, gpu, std::vector<int32_t>(prompt.begin() + chunk, prompt.end()), chunk, rec, true);
    } else {
        tok = step(model, *sampler, gpu, prompt, 0, rec, true);
    }
    for (std::size_t s = 0; s < n_decode; s++) tok = step(model, *sampler, gpu, {tok}, n_prompt + s, rec, true);
}

} // namespace

int
main(int argc, char** argv)
{
    if (argc < 6) {
        std::fprintf(stderr, "usage: ref_driver <llama|qlora|gemma> <small|hd128> <out.bin> <n_prompt> <n_decode> [chunk]\n");
        return 2;
    }
    const std::string kind = argv[1], shape = argv[2], out_path = argv[3];
    const std::size_t n_prompt = std::stoul(argv[4]), n_decode = std::stoul(argv[5]), chunk = argc > 6 ? std::stoul(argv[6]) : 0;
    const shape_cfg c = shape == "hd128" ? kHd128 : kSmall;
    try {
        hardware_accelerator gpu(64);
        std::fprintf(stderr, "ref_driver: %s on '%s'\n", kind.c_str(), gpu.name().c_str());
        recorder rec{c.vocab, {}, {}, {}};
        if (kind == "llama" || kind == "qlora") {
            nn::llama3_options o{.head_dim = c.head_dim, .n_heads = c.n_heads, .n_kv_heads = c.n_kv_heads, .n_layers = c.n_layers,
                                 .max_seq_len = c.max_seq, .rope_theta = 500000.0f, .norm_eps = 1e-5f};
            using Llama = nn::llama3<bf16>;
            nn::indirect_layer<Llama> model(o, gpu);
            const bool quant = kind == "qlora";
            if (quant) {
                // the reference's own layer swap: lora_linear(2.0, 32) / lora_embedding / quantization::linear for `output`
                huggingface::llama3_qlora_safetensor_serializer<bf16, Llama> serializer(o, gpu);
                serializer.adapt(model);
            }
            init_parameters(*model, c, quant, /*tie_output=*/!quant, gpu);
            run(*model, gpu, c, n_prompt, n_decode, chunk, rec);
        } else if (kind == "gemma") {
            nn::gemma3_options o{.head_dim = c.head_dim, .hidden_dim = c.dim, .n_heads = c.n_heads, .n_kv_heads = c.n_kv_heads, .n_layers = c.n_layers,
                                 .max_seq_len = c.max_seq, .sliding_window = 16, .sliding_stride = 3, .attn_scale = float(c.head_dim),
                                 .rope_theta = 1000000.0f, .rope_sliding_theta = 10000.0f, .norm_eps = 1e-6f};
            nn::indirect_layer<nn::gemma3<bf16>> model(o, gpu);
            init_parameters(*model, c, false, /*tie_output=*/true, gpu);
            run(*model, gpu, c, n_prompt, n_decode, chunk, rec);
        } else {
            throw std::invalid_argument("unknown model kind " + kind);
        }
        std::ofstream os(out_path, std::ios::binary);
        const std::uint32_t vocab = std::uint32_t(c.vocab), n = std::uint32_t(rec.sampled.size());
        os.write("MCRF", 4);
        os.write(reinterpret_cast<const char*>(&vocab), 4);
        os.write(reinterpret_cast<const char*>(&n), 4);
        os.write(reinterpret_cast<const char*>(rec.logits.data()), std::streamsize(rec.logits.size() * 2));
        os.write(reinterpret_cast<const char*>(rec.sampled.data()), std::streamsize(rec.sampled.size() * 4));
        os.write(reinterpret_cast<const char*>(rec.greedy.data()), std::streamsize(rec.greedy.size() * 4));
        std::fprintf(stderr, "ref_driver: %u forwards written to %s\n", n, out_path.c_str());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
