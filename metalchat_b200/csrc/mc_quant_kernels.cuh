// metalchat_b200/csrc/mc_quant_kernels.cuh — decode GEMV over the reference's quantised weights.
//
// Reference path (quantization/lora.h:94-122, quantization/linear.h:17-64, kernel/mul.metal:59-85): every call
// dequantises the WHOLE int8 matrix to bf16 with `hadamard_broadcast` (w = r(r(q) * r(s)): scale rounded to bf16,
// product rounded to bf16), runs bmm, then two LoRA bmms, a scalar_mul and an add — 6 launches and ~5 bytes of
// traffic per weight.  Here the weights stay packed in HBM and are dequantised in registers with exactly those two
// roundings (HSUB2/HMUL2 in bf16x2 are single-rounding), then fed to the tensor cores as the A operand of
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32).  The batch rows are the B operand, so up to 8 activation rows cost
// the same weight traffic as one.
//
// Packed layouts ("fragment order": what a lane loads with one 16-byte request is what its mma needs):
//   WF_W4     super-unit = 8 units = 16 weight rows; k-tile = 64.  [super][ktile][lane][16 B]; in each 32-bit word j
//             (k-block 16j..16j+15 of the tile) nibble i and nibble i+4 form the bf16x2 A register a_i:
//             a0 = (row g, k 2t,2t+1)  a1 = (row g+8, same k)  a2 = (row g, k 2t+8,2t+9)  a3 = (row g+8, same k)
//             with g = lane / 4, t = lane % 4 and nibble = q + 8.  Scales: bf16 r(s), [super][ktile][g][4] =
//             {row g grp 0, row g grp 1, row g+8 grp 0, row g+8 grp 1} (group = 32 k).
//   WF_W8ROW  int8 with one fp32 scale per row (tok_embeddings / output): [super][ktile32][lane][16 B], bytes
//             {a0.lo,a0.hi,a1.lo,a1.hi,a2.lo,a2.hi,a3.lo,a3.hi} of k-block 0 then of k-block 1.
// "unit" -> rows is the same mapping as the bf16 GEMV (rope pairs / gate-up pairs / adjacent rows).
#pragma once
#include "mc_decode_kernels.cuh"

namespace mc {

constexpr int kQMaxMB = 8;      // activation rows per pass = n dimension of the mma
constexpr int kQPad = 8;        // bf16 elements of padding per staged activation row (bank-conflict-free B fragments)

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// bf16x2 single-rounding arithmetic on packed registers
__device__ __forceinline__ uint32_t hsub2_bf16(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// byte `sel` (0..3) of w as a signed int8 -> fp32, exactly, without the conversion pipe: the byte (xor 0x80, i.e. +128)
// is permuted into the mantissa of 2^23, then (2^23 + 128) is subtracted
__device__ __forceinline__ float s8_to_f32(uint32_t w_xor80, uint32_t sel)
{
    return __uint_as_float(__byte_perm(w_xor80, 0x4b000000u, 0x7650u + sel)) - 8388736.0f;
}
// two fp32 -> packed bf16x2 (lo = first), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// (x & mask) | magic in ONE LOP3 (with immediate operands the compiler emits two); mask / magic are kept in registers by the caller
__device__ __forceinline__ uint32_t q_and_or(uint32_t x, uint32_t mask, uint32_t magic)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(x), "r"(mask), "r"(magic));
    return d;
}

struct qgemv_params {
    gemv_params g;            // shapes, activations, epilogue operands (g.W = packed weights, g.N rows, g.K)
    const void* scales;       // WF_W4: packed bf16 scales; WF_W8ROW: fp32 [N] per-row scales
    const uint16_t* lora_b;   // [N, rank] bf16 (row order = weight row order) or null
    const uint16_t* lora_ax;  // [rows, ax_ld] bf16: r(A . x) of the stacked adaptors
    uint32_t ax_ld;           // row pitch of lora_ax
    uint32_t ax_slices;       // 1, 2 (w1|w3 interleaved) or 3 (q|k|v): which 16-wide slice of ax a row uses
    uint32_t slice_rows0, slice_rows1; // q|k|v: rows < slice_rows0 -> slice 0, < slice_rows1 -> slice 1, else 2
    uint32_t rank;
    float lora_scale;         // r(scale) as fp32
};

// slice of the stacked LoRA-A output that weight row `r` multiplies
__device__ __forceinline__ uint32_t lora_slice(const qgemv_params& q, uint32_t r)
{
    if (q.ax_slices == 2) return r & 1u;
    if (q.ax_slices == 3) return r < q.slice_rows0 ? 0u : (r < q.slice_rows1 ? 1u : 2u);
    return 0u;
}

// grid: any; CTA = 8 warps = (8/KS) super-units x KS k-slices.
template <int FMT, int PRO, int EPI, int KS>
__global__ void __launch_bounds__(kGemvThreads, 2) gemv_q_kernel(const qgemv_params qp)
{
    const gemv_params& p = qp.g;
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t ldsx = p.K + kQPad;
    uint16_t* sx = reinterpret_cast<uint16_t*>(smem);                                  // [rows + 1][K + pad] bf16; the last row is zero
    float* sred = reinterpret_cast<float*>(smem + size_t(p.rows + 1) * ldsx * 2);     // [8 warps][32 lanes][4]
    float* sscr = sred + kGemvWarps * 32 * 4;                                          // [8]

    constexpr uint32_t KT = FMT == WF_W4 ? 64 : 32; // k per 16-byte lane load
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g = lane >> 2, t = lane & 3;
    constexpr uint32_t UPC = kGemvWarps / KS;
    const uint32_t slot = warp / KS, ks = warp % KS;
    const uint32_t supers = (p.N / 2 + 7) / 8;
    const uint32_t ktiles = p.K / KT;
    const uint32_t kt_per = ktiles / KS; // host guarantees divisibility
    const uint32_t kt_beg = ks * kt_per, kt_end = kt_beg + kt_per;
    const uint4* Wq = static_cast<const uint4*>(p.W);

    uint32_t su = blockIdx.x * UPC + slot;
    const uint32_t sstride = gridDim.x * UPC;

    // first weight chunk before the activations are touched (PDL: the producer may still be running)
    constexpr int U = 4;
    uint4 wv[U];
    uint2 sv[U]; // WF_W4: 4 bf16 scales per load
    if (su < supers) {
#pragma unroll
        for (int i = 0; i < U; i++) {
            const uint32_t kt = kt_beg + i;
            if (kt < kt_end) {
                wv[i] = ldg_stream(Wq + (size_t(su) * ktiles + kt) * 32 + lane);
                if (FMT == WF_W4) sv[i] = ldg_stream8(static_cast<const uint2*>(qp.scales) + (size_t(su) * ktiles + kt) * 8 + g);
            }
        }
    }
    pdl_launch_dependents();
    pdl_wait();
    unsigned tp_parity = 0;
    if (EPI == EPI_PARTIAL_TP) tp_parity = *reinterpret_cast<volatile unsigned*>(p.tp.epoch) & 1u; // the epoch this launch will publish is +1

    // stage activation rows (zero rows above p.rows: they are the unused columns of the mma)
    for (uint32_t m = 0; m <= p.rows; m++) {
        uint16_t* dst = sx + size_t(m) * ldsx;
        if (m >= p.rows) {
            for (uint32_t k = threadIdx.x * 8; k < p.K; k += kGemvThreads * 8) *reinterpret_cast<uint4*>(dst + k) = make_uint4(0, 0, 0, 0);
            continue;
        }
        const uint16_t* xr = p.x + size_t(m) * p.ldx;
        if (PRO == PRO_RMSNORM) {
            float part = 0.0f;
            for (uint32_t k = threadIdx.x * 8; k < p.K; k += kGemvThreads * 8) {
                const uint4 v = *reinterpret_cast<const uint4*>(xr + k);
                float f;
                f = bf_lo(v.x), part = fmaf(f, f, part);
                f = bf_hi(v.x), part = fmaf(f, f, part);
                f = bf_lo(v.y), part = fmaf(f, f, part);
                f = bf_hi(v.y), part = fmaf(f, f, part);
                f = bf_lo(v.z), part = fmaf(f, f, part);
                f = bf_hi(v.z), part = fmaf(f, f, part);
                f = bf_lo(v.w), part = fmaf(f, f, part);
                f = bf_hi(v.w), part = fmaf(f, f, part);
            }
            const float total = block_sum_256(part, sscr);
            const float inv = 1.0f / sqrtf(__fadd_rn(total / float(p.K), p.eps));
            for (uint32_t k = threadIdx.x * 8; k < p.K; k += kGemvThreads * 8) {
                const uint4 v = *reinterpret_cast<const uint4*>(xr + k);
                const uint4 gw = *reinterpret_cast<const uint4*>(p.norm_w + k);
                uint4 o;
#define MC_NORM2(dst, vv, gg)                                                                           \
    dst = uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_lo(gg), bf_lo(vv)), inv))) |                 \
          (uint32_t(f32_to_bf16_bits(__fmul_rn(__fmul_rn(bf_hi(gg), bf_hi(vv)), inv))) << 16)
                MC_NORM2(o.x, v.x, gw.x);
                MC_NORM2(o.y, v.y, gw.y);
                MC_NORM2(o.z, v.z, gw.z);
                MC_NORM2(o.w, v.w, gw.w);
#undef MC_NORM2
                *reinterpret_cast<uint4*>(dst + k) = o;
            }
        } else {
            for (uint32_t k = threadIdx.x * 8; k < p.K; k += kGemvThreads * 8)
                *reinterpret_cast<uint4*>(dst + k) = *reinterpret_cast<const uint4*>(xr + k);
        }
    }
    // this thread finalises unit (su*8 + g) for activation rows 2t and 2t+1
    const bool fin0 = ks == 0 && 2 * t < p.rows, fin1 = ks == 0 && 2 * t + 1 < p.rows;
    int32_t pos0 = 0, pos1 = 0, seq0 = 0, seq1 = 0;
    if (EPI == EPI_QKV) {
        if (fin0) pos0 = p.row_pos[2 * t], seq0 = p.row_seq[2 * t];
        if (fin1) pos1 = p.row_pos[2 * t + 1], seq1 = p.row_seq[2 * t + 1];
    }
    __syncthreads();

    const uint16_t* xb = sx + size_t(g < p.rows ? g : p.rows) * ldsx + 2 * t; // B fragment base of this lane (batch row g, or the zero row)
    for (; su - slot < supers; su += sstride) {
        const bool active = su < supers;
        const uint32_t unit = su * 8 + g;
        const bool unit_ok = active && unit < p.N / 2;
        uint32_t r0 = 0, r1 = 0;
        if (unit_ok) unit_rows(EPI, p, unit, r0, r1);
        // epilogue operands early: residual / rope table / LoRA-B rows
        float e00 = 0, e01 = 0, e10 = 0, e11 = 0; // [row r0|r1][batch 2t|2t+1]
        if (unit_ok && (fin0 || fin1)) {
            if (EPI == EPI_RESIDUAL) {
                if (fin0) e00 = bf16_bits_to_f32(p.res[size_t(2 * t) * p.ldy + r0]), e10 = bf16_bits_to_f32(p.res[size_t(2 * t) * p.ldy + r1]);
                if (fin1) e01 = bf16_bits_to_f32(p.res[size_t(2 * t + 1) * p.ldy + r0]), e11 = bf16_bits_to_f32(p.res[size_t(2 * t + 1) * p.ldy + r1]);
            } else if (EPI == EPI_QKV) {
                const uint32_t half = p.head_dim >> 1, j = r0 % p.head_dim;
                if (fin0) e00 = p.fcos[size_t(pos0) * half + j], e10 = p.fsin[size_t(pos0) * half + j];
                if (fin1) e01 = p.fcos[size_t(pos1) * half + j], e11 = p.fsin[size_t(pos1) * half + j];
            }
        }
        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t nib_mask = 0x000f000fu, nib_magic = 0x43004300u; // in registers (not folded back into immediates): one LOP3 per nibble pair
        asm volatile("" : "+r"(nib_mask), "+r"(nib_magic));
        const uint32_t sn = su + sstride;
        const bool next_active = sn < supers;
        float rs0 = 0.0f, rs1 = 0.0f; // WF_W8ROW: r(scale) of rows r0 / r1
        if (FMT == WF_W8ROW && unit_ok) {
            rs0 = rbf(static_cast<const float*>(qp.scales)[r0]);
            rs1 = rbf(static_cast<const float*>(qp.scales)[r1]);
        }
        if (active) {
            for (uint32_t kc = kt_beg; kc < kt_end; kc += U) {
                uint4 nw[U];
                uint2 ns[U];
                const uint32_t kn = kc + U;
                if (kn < kt_end) {
#pragma unroll
                    for (int i = 0; i < U; i++) {
                        const uint32_t kt = kn + i;
                        if (kt < kt_end) {
                            nw[i] = ldg_stream(Wq + (size_t(su) * ktiles + kt) * 32 + lane);
                            if (FMT == WF_W4) ns[i] = ldg_stream8(static_cast<const uint2*>(qp.scales) + (size_t(su) * ktiles + kt) * 8 + g);
                        }
                    }
                } else if (next_active) {
#pragma unroll
                    for (int i = 0; i < U; i++) {
                        const uint32_t kt = kt_beg + i;
                        if (kt < kt_end) {
                            nw[i] = ldg_stream(Wq + (size_t(sn) * ktiles + kt) * 32 + lane);
                            if (FMT == WF_W4) ns[i] = ldg_stream8(static_cast<const uint2*>(qp.scales) + (size_t(sn) * ktiles + kt) * 8 + g);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < U; i++) {
                    const uint32_t kt = kc + i;
                    if (kt < kt_end) {
                        const uint16_t* xk = xb + size_t(kt) * KT;
                        if (FMT == WF_W4) {
                            const uint32_t words[4] = {wv[i].x, wv[i].y, wv[i].z, wv[i].w};
                            // scales of (row g, row g+8) for k-groups 0 and 1 of this tile, as (s,s) bf16x2
                            const uint32_t s_g0 = __byte_perm(sv[i].x, 0, 0x1010), s_g1 = __byte_perm(sv[i].x, 0, 0x3232);
                            const uint32_t s_h0 = __byte_perm(sv[i].y, 0, 0x1010), s_h1 = __byte_perm(sv[i].y, 0, 0x3232);
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const uint32_t w = words[j];
                                const uint32_t sg = j < 2 ? s_g0 : s_g1, sh = j < 2 ? s_h0 : s_h1;
                                // nibble -> bf16 (128 + n) -> q = n - 8 (exact) -> r(q * r(s))   (kernel/mul.metal:76-77)
                                const uint32_t a0 = hmul2_bf16(hsub2_bf16(q_and_or(w, nib_mask, nib_magic), 0x43084308u), sg);
                                const uint32_t a1 = hmul2_bf16(hsub2_bf16(q_and_or(w >> 4, nib_mask, nib_magic), 0x43084308u), sh);
                                const uint32_t a2 = hmul2_bf16(hsub2_bf16(q_and_or(w >> 8, nib_mask, nib_magic), 0x43084308u), sg);
                                const uint32_t a3 = hmul2_bf16(hsub2_bf16(q_and_or(w >> 12, nib_mask, nib_magic), 0x43084308u), sh);
                                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(xk + j * 16);
                                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(xk + j * 16 + 8);
                                mma_bf16_16816(c, a0, a1, a2, a3, b0, b1);
                            }
                        } else {
                            const uint32_t words[4] = {wv[i].x, wv[i].y, wv[i].z, wv[i].w};
#pragma unroll
                            for (int j = 0; j < 2; j++) {
                                const uint32_t lo = words[2 * j] ^ 0x80808080u, hi = words[2 * j + 1] ^ 0x80808080u;
                                // int8 -> fp32 (exact) * r(s) (exact in fp32) -> one rounding to bf16
                                const uint32_t a0 = pack_bf16x2(__fmul_rn(s8_to_f32(lo, 0), rs0), __fmul_rn(s8_to_f32(lo, 1), rs0));
                                const uint32_t a1 = pack_bf16x2(__fmul_rn(s8_to_f32(lo, 2), rs1), __fmul_rn(s8_to_f32(lo, 3), rs1));
                                const uint32_t a2 = pack_bf16x2(__fmul_rn(s8_to_f32(hi, 0), rs0), __fmul_rn(s8_to_f32(hi, 1), rs0));
                                const uint32_t a3 = pack_bf16x2(__fmul_rn(s8_to_f32(hi, 2), rs1), __fmul_rn(s8_to_f32(hi, 3), rs1));
                                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(xk + j * 16);
                                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(xk + j * 16 + 8);
                                mma_bf16_16816(c, a0, a1, a2, a3, b0, b1);
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < U; i++) wv[i] = nw[i], sv[i] = ns[i];
            }
        }
        // join the k-slices of the super-unit
        if (KS > 1) {
            float4* sr = reinterpret_cast<float4*>(sred);
            sr[warp * 32 + lane] = make_float4(c[0], c[1], c[2], c[3]);
            __syncthreads();
            if (ks == 0) {
                float4 a = make_float4(0, 0, 0, 0);
#pragma unroll
                for (int s = 0; s < KS; s++) {
                    const float4 v = sr[(warp + s) * 32 + lane];
                    a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
                }
                c[0] = a.x, c[1] = a.y, c[2] = a.z, c[3] = a.w;
            }
            __syncthreads();
        }
        // epilogue: c[0],c[1] = row r0 for batch 2t, 2t+1; c[2],c[3] = row r1
        if (unit_ok) {
#pragma unroll
            for (int bsel = 0; bsel < 2; bsel++) {
                if (!(bsel == 0 ? fin0 : fin1)) continue;
                const uint32_t m = 2 * t + bsel;
                if (EPI == EPI_PARTIAL_TP) {
                    // row-parallel linear under tensor parallelism: the unrounded fp32 partial sums of this rank's k range go to its own slot of
                    // the exchange row; the adaptor term and the rounding follow the all-reduce (tp_finish_lora_kernel)
                    float* dst = tp_slot(p.tp, p.tp.peer_buf[p.tp.rank], tp_parity, p.tp.rank) + size_t(m) * p.tp.dim;
                    dst[r0] = bsel == 0 ? c[0] : c[1];
                    dst[r1] = bsel == 0 ? c[2] : c[3];
                    continue;
                }
                float y0 = rbf(bsel == 0 ? c[0] : c[1]), y1 = rbf(bsel == 0 ? c[2] : c[3]);
                if (qp.lora_b) {
                    // y = r(y + r(r(B . ax) * scale))   (quantization/lora.h:115-122); 16-byte loads, all issued before use
                    const uint16_t* ax0 = qp.lora_ax + size_t(m) * qp.ax_ld + lora_slice(qp, r0) * qp.rank;
                    const uint16_t* ax1 = qp.lora_ax + size_t(m) * qp.ax_ld + lora_slice(qp, r1) * qp.rank;
                    const uint16_t* b0p = qp.lora_b + size_t(r0) * qp.rank;
                    const uint16_t* b1p = qp.lora_b + size_t(r1) * qp.rank;
                    float l0 = 0.0f, l1 = 0.0f;
                    if ((qp.rank & 7u) == 0) {
                        for (uint32_t j = 0; j < qp.rank; j += 8) {
                            const uint4 xa = *reinterpret_cast<const uint4*>(ax0 + j), xb2 = *reinterpret_cast<const uint4*>(ax1 + j);
                            const uint4 wa = *reinterpret_cast<const uint4*>(b0p + j), wb = *reinterpret_cast<const uint4*>(b1p + j);
                            l0 = dot8(wa, xa, l0);
                            l1 = dot8(wb, xb2, l1);
                        }
                    } else {
                        for (uint32_t j = 0; j < qp.rank; j++) {
                            l0 = fmaf(bf16_bits_to_f32(ax0[j]), bf16_bits_to_f32(b0p[j]), l0);
                            l1 = fmaf(bf16_bits_to_f32(ax1[j]), bf16_bits_to_f32(b1p[j]), l1);
                        }
                    }
                    y0 = rbf(__fadd_rn(y0, rbf(__fmul_rn(rbf(l0), qp.lora_scale))));
                    y1 = rbf(__fadd_rn(y1, rbf(__fmul_rn(rbf(l1), qp.lora_scale))));
                }
                const float ea = bsel == 0 ? e00 : e01, eb = bsel == 0 ? e10 : e11;
                if (EPI == EPI_NONE) {
                    p.y[size_t(m) * p.ldy + r0] = f32_to_bf16_bits(y0);
                    p.y[size_t(m) * p.ldy + r1] = f32_to_bf16_bits(y1);
                } else if (EPI == EPI_RESIDUAL) {
                    p.y[size_t(m) * p.ldy + r0] = f32_to_bf16_bits(__fadd_rn(ea, y0));
                    p.y[size_t(m) * p.ldy + r1] = f32_to_bf16_bits(__fadd_rn(eb, y1));
                } else if (EPI == EPI_SWIGLU) {
                    p.y[size_t(m) * p.ldy + (r0 >> 1)] = f32_to_bf16_bits(__fmul_rn(silu_bf16(y0), y1));
                } else { // EPI_QKV
                    const uint32_t hd = p.head_dim, half = hd >> 1;
                    const uint32_t head = r0 / hd, j = r0 - head * hd;
                    const int32_t pos = bsel == 0 ? pos0 : pos1, seq = bsel == 0 ? seq0 : seq1;
                    if (head < p.n_heads + p.n_kv_heads) {
                        const float o0 = rbf(__fsub_rn(__fmul_rn(ea, y0), __fmul_rn(eb, y1)));
                        const float o1 = rbf(__fadd_rn(__fmul_rn(eb, y0), __fmul_rn(ea, y1)));
                        y0 = o0, y1 = o1;
                    }
                    if (head < p.n_heads) {
                        uint16_t* dst = p.q + size_t(m) * p.n_heads * hd + size_t(head) * hd + j;
                        dst[0] = f32_to_bf16_bits(y0);
                        dst[half] = f32_to_bf16_bits(y1);
                    } else {
                        const bool is_k = head < p.n_heads + p.n_kv_heads;
                        const uint32_t kvh = is_k ? head - p.n_heads : head - p.n_heads - p.n_kv_heads;
                        uint16_t* base = is_k ? p.kcache : p.vcache;
                        uint16_t* dst = base + ((size_t(seq) * p.n_kv_heads + kvh) * p.max_seq + size_t(min(uint32_t(pos), p.max_seq - 1))) * hd + j; // beyond the cache: the last row (sink roll, nn/cache.h:183-204)
                        dst[0] = f32_to_bf16_bits(y0);
                        dst[half] = f32_to_bf16_bits(y1);
                    }
                }
            }
        }
    }
    if (EPI == EPI_PARTIAL_TP) tp_publish(p.tp, p.rows, tp_parity, reinterpret_cast<unsigned*>(sscr));
}

// Tensor parallel QLoRA, after a row-parallel linear (wo / w2): every rank holds the `world` partial exchange rows
// [dim main sums | rank adaptor sums] and finishes ALL rows of the residual stream itself, in the reference's order of roundings
// (quantization/lora.h:115-122, nn/transformer.h:133,139):
//   y = r(sum_k main_k),  ax = r(sum_k ax_k),  y = r(y + r(r(B . ax) * r(scale))),  out = r(res + y)
// with the ranks summed in rank order.  One thread per output element; launched between the producer GEMV and the next phase.
struct tp_finish_params {
    tp_exchange tp;
    const uint16_t* res;      // [rows, ld]
    uint16_t* out;            // [rows, ld]
    const uint16_t* lora_b;   // [dim, rank] bf16
    uint32_t rows, dim, ld, rank;
    float lora_scale;         // r(scale) as fp32
};
MC_KERNEL void __launch_bounds__(256) tp_finish_lora_kernel(const tp_finish_params q)
{
    __shared__ unsigned word;
    __shared__ __align__(16) uint16_t sax[kMaxMB][64];
    pdl_launch_dependents();
    pdl_wait();
    const unsigned e = tp_wait(q.tp, &word);
    const uint32_t parity = (e - 1) & 1u;
    float* mine = q.tp.peer_buf[q.tp.rank];
    for (uint32_t i = threadIdx.x; i < q.rows * q.rank; i += blockDim.x) {
        const uint32_t m = i / q.rank, j = i - m * q.rank;
        float a = 0.0f;
        for (uint32_t src = 0; src < q.tp.world; src++) a += __ldcg(tp_slot(q.tp, mine, parity, src) + size_t(m) * q.tp.dim + q.dim + j);
        sax[m][j] = f32_to_bf16_bits(a);
    }
    __syncthreads();
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= q.dim) return;
    for (uint32_t m = 0; m < q.rows; m++) {
        float sum = 0.0f;
        for (uint32_t src = 0; src < q.tp.world; src++) sum += __ldcg(tp_slot(q.tp, mine, parity, src) + size_t(m) * q.tp.dim + k);
        float y = rbf(sum);
        const uint16_t* bp = q.lora_b + size_t(k) * q.rank;
        float l = 0.0f;
        if ((q.rank & 7u) == 0) {
            for (uint32_t j = 0; j < q.rank; j += 8) l = dot8(*reinterpret_cast<const uint4*>(bp + j), *reinterpret_cast<const uint4*>(&sax[m][j]), l);
        } else {
            for (uint32_t j = 0; j < q.rank; j++) l = fmaf(bf16_bits_to_f32(sax[m][j]), bf16_bits_to_f32(bp[j]), l);
        }
        y = rbf(__fadd_rn(y, rbf(__fmul_rn(rbf(l), q.lora_scale))));
        q.out[size_t(m) * q.ld + k] = f32_to_bf16_bits(__fadd_rn(bf16_bits_to_f32(q.res[size_t(m) * q.ld + k]), y));
    }
}

// ---- packing (bit-exact round trip; the reference stores one int4-range weight per int8, huggingface/llama.h:152-171) ----
// q8: int8 [N, K] row-major (values in [-8, 7]); epi selects the unit -> rows mapping of the consumer kernel.
MC_KERNEL void pack_w4_kernel(uint32_t* out, const int8_t* q8, gemv_params p, int epi, int* bad)
{
    const uint32_t ktiles = p.K / 64, supers = (p.N / 2 + 7) / 8;
    const uint64_t words = uint64_t(supers) * ktiles * 32 * 4;
    for (uint64_t w = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; w < words; w += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t j = uint32_t(w & 3), lane = uint32_t((w >> 2) & 31);
        const uint64_t tile = w >> 7;
        const uint32_t kt = uint32_t(tile % ktiles), su = uint32_t(tile / ktiles);
        const uint32_t g = lane >> 2, t = lane & 3;
        const uint32_t unit = su * 8 + g;
        uint32_t word = 0x88888888u; // q = 0 for rows beyond N
        if (unit < p.N / 2) {
            uint32_t r0, r1;
            unit_rows(epi, p, unit, r0, r1);
            const uint32_t kb = kt * 64 + j * 16;
            const uint32_t rr[8] = {r0, r1, r0, r1, r0, r1, r0, r1};
            const uint32_t kk[8] = {2 * t, 2 * t, 2 * t + 8, 2 * t + 8, 2 * t + 1, 2 * t + 1, 2 * t + 9, 2 * t + 9};
            word = 0;
            for (int n = 0; n < 8; n++) {
                const int v = q8[size_t(rr[n]) * p.K + kb + kk[n]];
                if (v < -8 || v > 7) atomicExch(bad, 1);
                word |= uint32_t((v + 8) & 15) << (4 * n);
            }
        }
        out[w] = word;
    }
}
MC_KERNEL void unpack_w4_kernel(int8_t* q8, const uint32_t* in, gemv_params p, int epi)
{
    const uint32_t ktiles = p.K / 64, supers = (p.N / 2 + 7) / 8;
    const uint64_t words = uint64_t(supers) * ktiles * 32 * 4;
    for (uint64_t w = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; w < words; w += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t j = uint32_t(w & 3), lane = uint32_t((w >> 2) & 31);
        const uint64_t tile = w >> 7;
        const uint32_t kt = uint32_t(tile % ktiles), su = uint32_t(tile / ktiles);
        const uint32_t g = lane >> 2, t = lane & 3;
        const uint32_t unit = su * 8 + g;
        if (unit >= p.N / 2) continue;
        uint32_t r0, r1;
        unit_rows(epi, p, unit, r0, r1);
        const uint32_t kb = kt * 64 + j * 16;
        const uint32_t rr[8] = {r0, r1, r0, r1, r0, r1, r0, r1};
        const uint32_t kk[8] = {2 * t, 2 * t, 2 * t + 8, 2 * t + 8, 2 * t + 1, 2 * t + 1, 2 * t + 9, 2 * t + 9};
        const uint32_t word = in[w];
        for (int n = 0; n < 8; n++) q8[size_t(rr[n]) * p.K + kb + kk[n]] = int8_t(int((word >> (4 * n)) & 15) - 8);
    }
}
// scales fp32 [N, K/32] -> bf16 r(s) in fragment order [super][ktile][g][4]
MC_KERNEL void pack_w4_scales_kernel(uint16_t* out, const float* s, gemv_params p, int epi)
{
    const uint32_t ktiles = p.K / 64, supers = (p.N / 2 + 7) / 8, groups = p.K / 32;
    const uint64_t n = uint64_t(supers) * ktiles * 8 * 4;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t e = uint32_t(i & 3), g = uint32_t((i >> 2) & 7);
        const uint64_t tile = i >> 5;
        const uint32_t kt = uint32_t(tile % ktiles), su = uint32_t(tile / ktiles);
        const uint32_t unit = su * 8 + g;
        uint16_t v = 0;
        if (unit < p.N / 2) {
            uint32_t r0, r1;
            unit_rows(epi, p, unit, r0, r1);
            const uint32_t row = e < 2 ? r0 : r1, grp = kt * 2 + (e & 1);
            v = f32_to_bf16_bits(s[size_t(row) * groups + grp]);
        }
        out[i] = v;
    }
}
// int8 [N, K] -> fragment order for WF_W8ROW (adjacent-rows units)
MC_KERNEL void pack_w8_kernel(uint32_t* out, const int8_t* q8, gemv_params p, int epi)
{
    const uint32_t ktiles = p.K / 32, supers = (p.N / 2 + 7) / 8;
    const uint64_t words = uint64_t(supers) * ktiles * 32 * 4;
    for (uint64_t w = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; w < words; w += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t j = uint32_t(w & 3), lane = uint32_t((w >> 2) & 31);
        const uint64_t tile = w >> 7;
        const uint32_t kt = uint32_t(tile % ktiles), su = uint32_t(tile / ktiles);
        const uint32_t g = lane >> 2, t = lane & 3;
        const uint32_t unit = su * 8 + g;
        uint32_t word = 0;
        if (unit < p.N / 2) {
            uint32_t r0, r1;
            unit_rows(epi, p, unit, r0, r1);
            const uint32_t kb = kt * 32 + (j >> 1) * 16 + (j & 1) * 8; // word 2j' = k 2t,2t+1 of rows r0,r1; word 2j'+1 = k 2t+8,2t+9
            const uint32_t k0 = kb + 2 * t;
            const uint8_t b0 = uint8_t(q8[size_t(r0) * p.K + k0]), b1 = uint8_t(q8[size_t(r0) * p.K + k0 + 1]);
            const uint8_t b2 = uint8_t(q8[size_t(r1) * p.K + k0]), b3 = uint8_t(q8[size_t(r1) * p.K + k0 + 1]);
            word = uint32_t(b0) | (uint32_t(b1) << 8) | (uint32_t(b2) << 16) | (uint32_t(b3) << 24);
        }
        out[w] = word;
    }
}

} // namespace mc
