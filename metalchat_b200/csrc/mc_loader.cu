// safetensors -> device loader.
//
// Replaces, for the decode path, safetensor_document::open / load of the reference (src/safetensor.cc:83-153,237-253;
// include/metalchat/safetensor.h:689-747): there the file is mapped and its pages are wrapped as no-copy Metal buffers (unified
// memory); a B200 has no unified memory, so here every tensor goes from the mapping to its device layout with one strided H2D
// copy per shard slice (mc_llama_set_tensor: column / row / vocabulary
// slices under tensor parallelism, fused wqkv / w13 rows, int8 -> packed int4 at finalize).  Name handling follows the
// reference's serializers: HuggingFace names are renamed to the registered layer paths (huggingface/llama.h:88-103), the output
// projection stays tied to the embedding unless the file carries its own (huggingface/llama.h:103, reference.h:53-59), and
// Meta-format checkpoints get the rows of wq / wk permuted [head, hd/2, 2] -> [head, 2, hd/2] (reference.h:73-94,
// nn/attention.h:225-255).  File format (src/safetensor.cc:83-133): u64 little-endian header length, JSON header
// {name: {dtype, shape, data_offsets}, "__metadata__": {...}}, raw little-endian data.
#include "mc_common.cuh"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <filesystem>
#include <map>
#include <memory>
#include <regex>
#include <string>
#include <vector>

using namespace mc;

namespace {

struct st_entry {
    std::string name, dtype;
    std::vector<uint64_t> shape;
    const char* data = nullptr;
    uint64_t nbytes = 0;
    uint32_t file = 0;
};

struct st_file {
    std::string path;
    void* map = nullptr;
    size_t size = 0;
};

// ---- a small JSON reader: exactly what a safetensors header needs (objects, arrays, strings, numbers, literals) -------------
struct json_reader {
    const char* p;
    const char* end;
    const std::string& where;
    [[noreturn]] void bad(const char* what) const { throw error(MC_ERR_INVALID, "safetensors: " + where + ": malformed header (" + what + ")"); }
    void ws()
    {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool eat(char c)
    {
        ws();
        if (p < end && *p == c) {
            p++;
            return true;
        }
        return false;
    }
    void expect(char c)
    {
        if (!eat(c)) bad("unexpected character");
    }
    static void utf8(std::string& out, uint32_t cp)
    {
        if (cp < 0x80) out += char(cp);
        else if (cp < 0x800) out += char(0xC0 | (cp >> 6)), out += char(0x80 | (cp & 0x3F));
        else if (cp < 0x10000) out += char(0xE0 | (cp >> 12)), out += char(0x80 | ((cp >> 6) & 0x3F)), out += char(0x80 | (cp & 0x3F));
        else out += char(0xF0 | (cp >> 18)), out += char(0x80 | ((cp >> 12) & 0x3F)), out += char(0x80 | ((cp >> 6) & 0x3F)), out += char(0x80 | (cp & 0x3F));
    }
    uint32_t hex4()
    {
        if (end - p < 4) bad("truncated \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            const char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
            else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
            else bad("bad \\u escape");
        }
        return v;
    }
    std::string string()
    {
        ws();
        if (p >= end || *p != '"') bad("string expected");
        p++;
        std::string out;
        while (true) {
            if (p >= end) bad("unterminated string");
            const char c = *p++;
            if (c == '"') return out;
            if (c != '\\') {
                // JSON strings hold no raw control characters and are valid UTF-8; names travel as C strings afterwards
                const unsigned char u = static_cast<unsigned char>(c);
                if (u < 0x20) bad("control character in a string");
                if (u >= 0x80) {
                    const int more = (u & 0xE0) == 0xC0 ? 1 : (u & 0xF0) == 0xE0 ? 2 : (u & 0xF8) == 0xF0 ? 3 : -1;
                    if (more < 0 || end - p < more || (u == 0xC0 || u == 0xC1) || u > 0xF4) bad("invalid UTF-8 in a string");
                    out += c;
                    for (int i = 0; i < more; i++) {
                        if ((static_cast<unsigned char>(*p) & 0xC0) != 0x80) bad("invalid UTF-8 in a string");
                        out += *p++;
                    }
                    continue;
                }
                out += c;
                continue;
            }
            if (p >= end) bad("unterminated escape");
            const char e = *p++;
            switch (e) {
            case '"': out += '"'; break;
            case '\\': out += '\\'; break;
            case '/': out += '/'; break;
            case 'b': out += '\b'; break;
            case 'f': out += '\f'; break;
            case 'n': out += '\n'; break;
            case 'r': out += '\r'; break;
            case 't': out += '\t'; break;
            case 'u': {
                uint32_t cp = hex4();
                if (cp == 0) bad("\\u0000 in a string");
                if (cp >= 0xD800 && cp < 0xDC00) {
                    // a high surrogate is only valid as the first half of a pair
                    if (!(end - p >= 6 && p[0] == '\\' && p[1] == 'u')) bad("lone surrogate in a string");
                    p += 2;
                    const uint32_t lo = hex4();
                    if (lo < 0xDC00 || lo >= 0xE000) bad("lone surrogate in a string");
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                } else if (cp >= 0xDC00 && cp < 0xE000) bad("lone surrogate in a string");
                utf8(out, cp);
                break;
            }
            default: bad("bad escape");
            }
        }
    }
    uint64_t u64()
    {
        ws();
        if (p >= end || *p < '0' || *p > '9') bad("non-negative integer expected");
        uint64_t v = 0;
        while (p < end && *p >= '0' && *p <= '9') {
            if (v > (UINT64_MAX - 9) / 10) bad("integer overflow");
            v = v * 10 + uint64_t(*p++ - '0');
        }
        if (p < end && (*p == '.' || *p == 'e' || *p == 'E')) bad("integer expected");
        return v;
    }
    void skip_value(int depth = 0)
    {
        if (depth > 64) bad("nesting too deep");
        ws();
        if (p >= end) bad("truncated");
        if (*p == '"') {
            string();
        } else if (*p == '{') {
            p++;
            if (eat('}')) return;
            do {
                string();
                expect(':');
                skip_value(depth + 1);
            } while (eat(','));
            expect('}');
        } else if (*p == '[') {
            p++;
            if (eat(']')) return;
            do skip_value(depth + 1);
            while (eat(','));
            expect(']');
        } else {
            const char* q = p;
            while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;
            if (p == q) bad("value expected");
        }
    }
};

// element size of a safetensors dtype tag (safetensor.h:251-264); 0 = unknown
uint32_t dtype_size(const std::string& d)
{
    static const std::map<std::string, uint32_t> sizes = {{"BOOL", 1}, {"I8", 1},  {"U8", 1},  {"I16", 2}, {"U16", 2}, {"F16", 2}, {"BF16", 2},
                                                          {"I32", 4},  {"U32", 4}, {"F32", 4}, {"F64", 8}, {"I64", 8}, {"U64", 8}};
    const auto it = sizes.find(d);
    return it == sizes.end() ? 0u : it->second;
}

} // namespace

struct mc_safetensors {
    std::vector<st_file> files;
    std::vector<st_entry> entries;
    std::map<std::string, uint32_t> by_name;
    std::map<std::string, std::string> metadata;

    ~mc_safetensors()
    {
        for (st_file& f : files) {
            if (f.map) munmap(f.map, f.size);
        }
    }
    void add_file(const std::string& path)
    {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw error(MC_ERR_INVALID, "safetensors: cannot open " + path);
        struct stat sb {};
        if (fstat(fd, &sb) != 0 || sb.st_size < 8) {
            ::close(fd);
            throw error(MC_ERR_INVALID, "safetensors: " + path + ": file is shorter than its 8-byte header length");
        }
        st_file f;
        f.path = path, f.size = size_t(sb.st_size);
        f.map = mmap(nullptr, f.size, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (f.map == MAP_FAILED) throw error(MC_ERR_RUNTIME, "safetensors: mmap failed for " + path);
        files.push_back(f);
        const uint32_t fi = uint32_t(files.size() - 1);
        const char* base = static_cast<const char*>(f.map);
        uint64_t hlen = 0;
        memcpy(&hlen, base, 8); // little-endian host
        if (hlen > f.size - 8) throw error(MC_ERR_INVALID, "safetensors: " + path + ": header length " + std::to_string(hlen) + " exceeds the file");
        const char* data0 = base + 8 + hlen;
        const uint64_t data_bytes = f.size - 8 - hlen;
        json_reader r{base + 8, base + 8 + hlen, path};
        r.expect('{');
        if (!r.eat('}')) {
            do {
                const std::string name = r.string();
                r.expect(':');
                if (name == "__metadata__") {
                    r.expect('{');
                    if (!r.eat('}')) {
                        do {
                            const std::string k = r.string();
                            r.expect(':');
                            r.ws();
                            if (r.p < r.end && *r.p == '"') metadata[k] = r.string();
                            else r.skip_value();
                        } while (r.eat(','));
                        r.expect('}');
                    }
                    continue;
                }
                st_entry e;
                e.name = name, e.file = fi;
                uint64_t off[2] = {0, 0};
                bool have_dtype = false, have_shape = false, have_off = false;
                r.expect('{');
                if (!r.eat('}')) {
                    do {
                        const std::string k = r.string();
                        r.expect(':');
                        if (k == "dtype") e.dtype = r.string(), have_dtype = true;
                        else if (k == "shape") {
                            r.expect('[');
                            if (!r.eat(']')) {
                                do e.shape.push_back(r.u64());
                                while (r.eat(','));
                                r.expect(']');
                            }
                            have_shape = true;
                        } else if (k == "data_offsets") {
                            r.expect('[');
                            off[0] = r.u64();
                            r.expect(',');
                            off[1] = r.u64();
                            r.expect(']');
                            have_off = true;
                        } else r.skip_value();
                    } while (r.eat(','));
                    r.expect('}');
                }
                if (!have_dtype || !have_shape || !have_off) throw error(MC_ERR_INVALID, "safetensors: " + path + ": tensor " + name + " lacks dtype / shape / data_offsets");
                const uint32_t es = dtype_size(e.dtype);
                if (es == 0) throw error(MC_ERR_INVALID, "safetensors: " + path + ": tensor " + name + " has unknown dtype " + e.dtype);
                uint64_t numel = 1;
                for (uint64_t d : e.shape) {
                    if (d != 0 && numel > UINT64_MAX / d) throw error(MC_ERR_INVALID, "safetensors: " + path + ": tensor " + name + ": shape overflows");
                    numel *= d;
                }
                if (off[0] > off[1] || off[1] > data_bytes || off[1] - off[0] != numel * es)
                    throw error(MC_ERR_INVALID, "safetensors: " + path + ": tensor " + name + ": data_offsets [" + std::to_string(off[0]) + ", " + std::to_string(off[1]) +
                                                    ") do not match " + std::to_string(numel) + " x " + e.dtype + " inside " + std::to_string(data_bytes) + " data bytes");
                e.data = data0 + off[0], e.nbytes = off[1] - off[0];
                if (by_name.count(name)) throw error(MC_ERR_INVALID, "safetensors: tensor " + name + " appears twice (" + path + ")");
                by_name[name] = uint32_t(entries.size());
                entries.push_back(std::move(e));
            } while (r.eat(','));
            r.expect('}');
        }
        r.ws();
        if (r.p != r.end) r.bad("trailing bytes");
    }
};

namespace {

// HuggingFace -> registered layer paths (huggingface/llama.h:88-103); lm_head is this loader's addition for checkpoints with an
// untied output projection (Llama-3.1-8B / 70B): the reference always aliases the embedding
std::string hf_to_meta(const std::string& name)
{
    static const std::vector<std::pair<std::regex, std::string>> mapping = {
        {std::regex(R"(model\.(layers\.\d+)\.input_layernorm)"), "$1.attention_norm"},
        {std::regex(R"(model\.(layers\.\d+)\.post_attention_layernorm)"), "$1.ffn_norm"},
        {std::regex(R"(model\.(layers\.\d+)\.mlp\.gate_proj)"), "$1.feed_forward.w1"},
        {std::regex(R"(model\.(layers\.\d+)\.mlp\.down_proj)"), "$1.feed_forward.w2"},
        {std::regex(R"(model\.(layers\.\d+)\.mlp\.up_proj)"), "$1.feed_forward.w3"},
        {std::regex(R"(model\.(layers\.\d+)\.self_attn\.q_proj)"), "$1.attention.wq"},
        {std::regex(R"(model\.(layers\.\d+)\.self_attn\.k_proj)"), "$1.attention.wk"},
        {std::regex(R"(model\.(layers\.\d+)\.self_attn\.v_proj)"), "$1.attention.wv"},
        {std::regex(R"(model\.(layers\.\d+)\.self_attn\.o_proj)"), "$1.attention.wo"},
        {std::regex(R"(model\.norm)"), "norm"},
        {std::regex(R"(model\.embed_tokens)"), "tok_embeddings"},
        {std::regex(R"(lm_head)"), "output"},
    };
    for (const auto& [re, to] : mapping) {
        std::smatch mt;
        if (std::regex_search(name, mt, re) && mt.position(0) == 0) return std::regex_replace(name, re, to, std::regex_constants::format_first_only);
    }
    return name;
}

// the parameters a model of this configuration registers with their element sizes
struct param_spec {
    std::string name;
    uint32_t elem;
    bool optional;
};
std::vector<param_spec> expected_params(const mc_llama_config& c)
{
    std::vector<param_spec> out;
    const bool Q = c.quant != 0;
    auto linear = [&](const std::string& p) {
        if (!Q) {
            out.push_back({p + ".weight", 2, false});
        } else {
            out.push_back({p + ".weight", 1, false});
            out.push_back({p + ".scales", 4, false});
            out.push_back({p + ".adaptor.A.weight", 2, false});
            out.push_back({p + ".adaptor.B.weight", 2, false});
        }
    };
    for (uint32_t i = 0; i < c.n_layers; i++) {
        const std::string p = "layers." + std::to_string(i) + ".";
        out.push_back({p + "attention_norm.weight", 2, false});
        out.push_back({p + "ffn_norm.weight", 2, false});
        for (const char* w : {"attention.wq", "attention.wk", "attention.wv", "attention.wo", "feed_forward.w1", "feed_forward.w2", "feed_forward.w3"}) linear(p + w);
    }
    out.push_back({"norm.weight", 2, false});
    if (!Q) {
        out.push_back({"tok_embeddings.weight", 2, false});
        out.push_back({"output.weight", 2, true}); // tied to the embedding when absent
    } else {
        out.push_back({"tok_embeddings.weight", 1, false});
        out.push_back({"tok_embeddings.scales", 4, false});
        out.push_back({"output.weight", 1, false});
        out.push_back({"output.scales", 4, false});
    }
    return out;
}

uint16_t f32_bits_to_bf16(uint32_t u)
{
    if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40u);
    return uint16_t((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
uint32_t f16_bits_to_f32(uint16_t h)
{
    const uint32_t sign = uint32_t(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
    if (exp == 0) {
        if (man == 0) return sign;
        int e = -1;
        do {
            man <<= 1;
            e++;
        } while (!(man & 0x400u));
        return sign | (uint32_t(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
    if (exp == 31) return sign | 0x7f800000u | (man << 13);
    return sign | ((exp + 112u) << 23) | (man << 13);
}

} // namespace

extern "C" {

mc_status mc_safetensors_open(const char* path, mc_safetensors** out)
{
    MC_API_BEGIN
    MC_REQUIRE(path && out, "bad arguments");
    namespace fs = std::filesystem;
    auto st = std::make_unique<mc_safetensors>();
    std::error_code ec;
    if (fs::is_directory(path, ec)) {
        // a sharded checkpoint: every *.safetensors file of the directory, in name order (safetensor.h:949-1036 reads the index;
        // the union of the files is the same set of tensors)
        std::vector<std::string> names;
        for (const auto& de : fs::directory_iterator(path))
            if (de.is_regular_file() && de.path().extension() == ".safetensors") names.push_back(de.path().string());
        std::sort(names.begin(), names.end());
        if (names.empty()) throw error(MC_ERR_INVALID, std::string("safetensors: no *.safetensors file in ") + path);
        for (const std::string& n : names) st->add_file(n);
    } else {
        st->add_file(path);
    }
    *out = st.release();
    MC_API_END
}

mc_status mc_safetensors_close(mc_safetensors* st)
{
    MC_API_BEGIN
    delete st;
    MC_API_END
}

mc_status mc_safetensors_count(mc_safetensors* st, uint32_t* n)
{
    MC_API_BEGIN
    MC_REQUIRE(st && n, "bad arguments");
    *n = uint32_t(st->entries.size());
    MC_API_END
}

static void fill_entry(const st_entry& e, mc_safetensors_entry* out)
{
    memset(out, 0, sizeof(*out));
    out->name = e.name.c_str(), out->dtype = e.dtype.c_str();
    out->rank = uint32_t(std::min<size_t>(e.shape.size(), 8));
    for (uint32_t i = 0; i < out->rank; i++) out->shape[i] = e.shape[i];
    out->data = e.data, out->nbytes = e.nbytes;
}

mc_status mc_safetensors_entry_at(mc_safetensors* st, uint32_t index, mc_safetensors_entry* out)
{
    MC_API_BEGIN
    MC_REQUIRE(st && out, "bad arguments");
    MC_REQUIRE(index < st->entries.size(), "safetensors: entry index out of range");
    MC_REQUIRE(st->entries[index].shape.size() <= 8, "safetensors: tensors of more than 8 dimensions are not supported");
    fill_entry(st->entries[index], out);
    MC_API_END
}

mc_status mc_safetensors_find(mc_safetensors* st, const char* name, mc_safetensors_entry* out)
{
    MC_API_BEGIN
    MC_REQUIRE(st && name && out, "bad arguments");
    const auto it = st->by_name.find(name);
    if (it == st->by_name.end()) throw error(MC_ERR_INVALID, std::string("safetensors: no tensor named ") + name);
    fill_entry(st->entries[it->second], out);
    MC_API_END
}

mc_status mc_safetensors_metadata(mc_safetensors* st, const char* key, const char** value)
{
    MC_API_BEGIN
    MC_REQUIRE(st && key && value, "bad arguments");
    const auto it = st->metadata.find(key);
    *value = it == st->metadata.end() ? nullptr : it->second.c_str();
    MC_API_END
}

mc_status mc_llama_load_safetensors(mc_llama* m, mc_safetensors* st, uint32_t flags, uint32_t* n_loaded)
{
    MC_API_BEGIN
    nvtx_range nvtx_("mc_llama_load_safetensors");
    MC_REQUIRE(m && st, "bad arguments");
    mc_llama_config cfg{};
    if (mc_llama_get_config(m, &cfg) != MC_OK) throw error(MC_ERR_RUNTIME, mc_last_error());
    // The H2D copies read the mapping directly (pageable source: the driver stages it, 11.7 GB/s on the B200 box).  Page-locking the
    // mapping in place was measured and dropped (tools/hostreg_probe.py): a read-only registration is refused on this platform, and a
    // private writable mapping registers at 0.5 ms per MiB (the pin breaks copy-on-write: 0.13 s for 256 MiB) -- the 55 GB/s copy that
    // follows does not earn that back for a tensor that is read once.
    std::map<std::string, const st_entry*> have;
    for (const st_entry& e : st->entries) {
        const std::string name = (flags & MC_LOAD_HF_NAMES) ? hf_to_meta(e.name) : e.name;
        if (have.count(name)) throw error(MC_ERR_INVALID, "safetensors: " + e.name + " and " + have[name]->name + " both map to parameter " + name);
        have[name] = &e;
    }
    uint32_t loaded = 0;
    std::vector<uint16_t> tmp;
    const uint32_t hd = cfg.head_dim;
    for (const param_spec& ps : expected_params(cfg)) {
        const auto it = have.find(ps.name);
        if (it == have.end()) {
            if (ps.optional || !(flags & MC_LOAD_STRICT)) continue;
            throw error(MC_ERR_INVALID, "safetensors: parameter " + ps.name + " is missing");
        }
        const st_entry& e = *it->second;
        const void* src = e.data;
        size_t nbytes = e.nbytes;
        const uint32_t es = dtype_size(e.dtype);
        const bool native = (ps.elem == 2 && e.dtype == "BF16") || (ps.elem == 1 && e.dtype == "I8") || (ps.elem == 4 && e.dtype == "F32");
        if (!native) {
            // fp32 / fp16 checkpoints of a bf16 model: one round-to-nearest-even on the host
            if (!(ps.elem == 2 && (e.dtype == "F32" || e.dtype == "F16")))
                throw error(MC_ERR_INVALID, "safetensors: parameter " + ps.name + " is " + e.dtype + ", the model stores " + (ps.elem == 2 ? "BF16" : ps.elem == 1 ? "I8" : "F32"));
            const size_t n = e.nbytes / es;
            tmp.resize(n);
            if (e.dtype == "F32") {
                const uint32_t* s = reinterpret_cast<const uint32_t*>(e.data);
                for (size_t i = 0; i < n; i++) tmp[i] = f32_bits_to_bf16(s[i]);
            } else {
                const uint16_t* s = reinterpret_cast<const uint16_t*>(e.data);
                for (size_t i = 0; i < n; i++) tmp[i] = f32_bits_to_bf16(f16_bits_to_f32(s[i]));
            }
            src = tmp.data(), nbytes = n * 2;
        }
        const bool is_q = ps.name.size() > 20 && ps.name.compare(ps.name.size() - 20, 20, ".attention.wq.weight") == 0;
        const bool is_k = ps.name.size() > 20 && ps.name.compare(ps.name.size() - 20, 20, ".attention.wk.weight") == 0;
        std::vector<char> perm;
        if ((flags & MC_LOAD_META_PERMUTE) && (is_q || is_k)) {
            // out[head*hd + k*hd/2 + j] = in[head*hd + 2j + k]   (nn/attention.h:232-247)
            if (e.shape.size() != 2 || e.shape[0] % hd != 0) throw error(MC_ERR_INVALID, "safetensors: " + ps.name + ": cannot permute heads of this shape");
            const size_t rows = e.shape[0], row_bytes = nbytes / rows;
            perm.resize(nbytes);
            const char* s = static_cast<const char*>(src);
            for (size_t r = 0; r < rows; r++) {
                const size_t head = r / hd, rem = r % hd, j = rem / 2, k = rem % 2;
                memcpy(perm.data() + (head * hd + k * (hd / 2) + j) * row_bytes, s + r * row_bytes, row_bytes);
            }
            src = perm.data();
        }
        if (mc_llama_set_tensor(m, ps.name.c_str(), src, nbytes) != MC_OK) throw error(MC_ERR_INVALID, std::string("safetensors: ") + e.name + ": " + mc_last_error());
        loaded++;
    }
    if (n_loaded) *n_loaded = loaded;
    MC_API_END
}

} // extern "C"
