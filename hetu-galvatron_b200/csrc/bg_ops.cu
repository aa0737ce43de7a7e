// bg_ops.cu -- local fused elementwise / row kernels that sit next to the collectives on the layer path
// (SURVEY 2.3 rows K5 RMSNorm, K6 swiglu, K9 RoPE + the split/transposes of transformer.py:731-867,
//  a10 vocab-parallel cross-entropy).  All HBM-bound: 16-B vector access, fp32 math, one rounding to bf16.
#include "bg_common.cuh"

using namespace bg;

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide all-reduce through shared memory; safe to call repeatedly
template <bool kMax>
__device__ __forceinline__ float block_reduce(float v, float* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = kMax ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    float r = (lane < nwarp) ? smem[lane] : (kMax ? -INFINITY : 0.f);
    r = kMax ? warp_max(r) : warp_sum(r);
    return r;
}

int local_grid(size_t items, int threads) {
    long long want = (long long)((items + threads - 1) / threads);
    if (want < 1) want = 1;
    return (int)(want < g_tun.local_ctas ? want : g_tun.local_ctas);
}

// ---------------------------------------------------------------------------------------------
// cast / scale / accumulate:  dst = [dst +] src * scale
// ---------------------------------------------------------------------------------------------
template <bool kSrcBf16, bool kDstBf16>
__global__ void __launch_bounds__(kThreads) cast_kernel(const void* __restrict__ src, void* __restrict__ dst, size_t nvec8,
                                                        float scale, int accumulate) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec8; v += stride) {
        float f[8];
        if (kSrcBf16) {
            unpack8(ld16_stream(reinterpret_cast<const uint4*>(src) + v), f);
        } else {
            uint4 a = ld16_stream(reinterpret_cast<const uint4*>(src) + 2 * v), b = ld16_stream(reinterpret_cast<const uint4*>(src) + 2 * v + 1);
            f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
            f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] *= scale;
        if (kDstBf16) {
            uint4* d = reinterpret_cast<uint4*>(dst) + v;
            if (accumulate) {
                float o[8];
                unpack8(*d, o);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += o[i];
            }
            st16(d, pack8(f));
        } else {
            float4* d = reinterpret_cast<float4*>(dst) + 2 * v;
            float4 lo = make_float4(f[0], f[1], f[2], f[3]), hi = make_float4(f[4], f[5], f[6], f[7]);
            if (accumulate) {
                float4 a = d[0], b = d[1];
                lo.x += a.x; lo.y += a.y; lo.z += a.z; lo.w += a.w;
                hi.x += b.x; hi.y += b.y; hi.z += b.z; hi.w += b.w;
            }
            d[0] = lo; d[1] = hi;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RMSNorm (flash_attn.ops.rms_norm semantics: fp32 math, y = x * rstd * w, one rounding)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxVpt = 4;  // register-resident row up to 256 * 8 * 4 = 8192 columns

// VPT = 16-B vectors per thread (1, 2 or 4: the row stays in registers); the NEXT row of the CTA is requested before the current
// one is reduced, so the loads of a CTA never drain while it sits in the two block-wide barriers.
template <int VPT>
__global__ void __launch_bounds__(kThreads) rmsnorm_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w,
                                                               uint4* __restrict__ y, float* __restrict__ rstd_out,
                                                               long long rows, int nvec, float eps) {
    __shared__ float smem[32];
    uint4 cur[VPT], nxt[VPT];
    long long r = blockIdx.x;
    if (r < rows) {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int v = threadIdx.x + j * kThreads;
            if (v < nvec) cur[j] = ld16_stream(x + r * nvec + v);
        }
    }
    for (; r < rows; r += gridDim.x) {
        const long long rn = r + gridDim.x;
        if (rn < rows) {
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const int v = threadIdx.x + j * kThreads;
                if (v < nvec) nxt[j] = ld16_stream(x + rn * nvec + v);
            }
        }
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int v = threadIdx.x + j * kThreads;
            if (v < nvec) {
                float f[8];
                unpack8(cur[j], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
            }
        }
        ss = block_reduce<false>(ss, smem);
        const float rstd = rsqrtf(ss / (float)(nvec * 8) + eps);
        if (threadIdx.x == 0 && rstd_out) rstd_out[r] = rstd;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int v = threadIdx.x + j * kThreads;
            if (v < nvec) {
                float f[8], g[8];
                unpack8(cur[j], f);
                unpack8(__ldg(w + v), g);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = f[i] * rstd * g[i];
                st16(y + r * nvec + v, pack8(f));
            }
        }
#pragma unroll
        for (int j = 0; j < VPT; ++j) cur[j] = nxt[j];
    }
}

// dx = rstd * (dy*w - xhat * mean(dy*w*xhat)),  dw_partial[cta] = sum over the CTA's rows of dy * xhat
template <int VPT>
__global__ void __launch_bounds__(kThreads) rmsnorm_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                               const uint4* __restrict__ w, const float* __restrict__ rstd_in,
                                                               uint4* __restrict__ dx, float* __restrict__ dw_partial,
                                                               long long rows, int nvec) {
    __shared__ float smem[32];
    float dw[VPT][8];
#pragma unroll
    for (int j = 0; j < VPT; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) dw[j][i] = 0.f;
    uint4 cx[VPT], cg[VPT], nx[VPT], ng[VPT];
    long long r = blockIdx.x;
    if (r < rows) {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int v = threadIdx.x + j * kThreads;
            if (v < nvec) { cx[j] = ld16_stream(x + r * nvec + v); cg[j] = ld16_stream(dy + r * nvec + v); }
        }
    }
    for (; r < rows; r += gridDim.x) {
        const long long rn = r + gridDim.x;
        if (rn < rows) {
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const int v = threadIdx.x + j * kThreads;
                if (v < nvec) { nx[j] = ld16_stream(x + rn * nvec + v); ng[j] = ld16_stream(dy + rn * nvec + v); }
            }
        }
        const float rstd = rstd_in[r];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int v = threadIdx.x + j * kThreads;
            if (v < nvec) {
                float xh[8], g[8], wv[8];
                unpack8(cx[j], xh);
                unpack8(cg[j], g);
                unpack8(__ldg(w + v), wv);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    xh[i] *= rstd;
                    dw[j][i] += g[i] * xh[i];
                    dot += g[i] * wv[i] * xh[i];
                }
            }
        }
        dot = block_reduce<false>(dot, smem) / (float)(nvec * 8);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {     // second pass over the packed registers (cheaper than keeping xhat and g*w unpacked)
            const int v = threadIdx.x + j * kThreads;
            if (v < nvec) {
                float xh[8], g[8], wv[8], o[8];
                unpack8(cx[j], xh);
                unpack8(cg[j], g);
                unpack8(__ldg(w + v), wv);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = rstd * (g[i] * wv[i] - xh[i] * rstd * dot);
                st16(dx + r * nvec + v, pack8(o));
            }
        }
#pragma unroll
        for (int j = 0; j < VPT; ++j) { cx[j] = nx[j]; cg[j] = ng[j]; }
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const int v = threadIdx.x + j * kThreads;
        if (v < nvec) {
            float4* d = reinterpret_cast<float4*>(dw_partial + ((size_t)blockIdx.x * nvec + v) * 8);
            d[0] = make_float4(dw[j][0], dw[j][1], dw[j][2], dw[j][3]);
            d[1] = make_float4(dw[j][4], dw[j][5], dw[j][6], dw[j][7]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm with bias (GPT / BERT families: torch.nn.LayerNorm in gpt_hf/GPTModel_tensor_parallel.py:34,56 and
// bert_hf/BertModel_tensor_parallel.py): fp32 math, y = (x - mean) * rstd * w + b, one rounding.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) layernorm_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w,
                                                                 const uint4* __restrict__ b, uint4* __restrict__ y,
                                                                 float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                 long long rows, int nvec, float eps) {
    __shared__ float smem[32];
    const float inv_n = 1.f / (float)(nvec * 8);
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const uint4* xr = x + r * nvec;
        uint4 xv[kMaxVpt];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVpt; ++j) {
            int v = threadIdx.x + j * kThreads;
            if (v < nvec) {
                xv[j] = ld16_stream(xr + v);
                float f[8];
                unpack8(xv[j], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) s += f[i];
            }
        }
        const float mean = block_reduce<false>(s, smem) * inv_n;
        float ss = 0.f;               // two-pass variance on the register-resident row (no cancellation)
#pragma unroll
        for (int j = 0; j < kMaxVpt; ++j) {
            int v = threadIdx.x + j * kThreads;
            if (v < nvec) {
                float f[8];
                unpack8(xv[j], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) ss += (f[i] - mean) * (f[i] - mean);
            }
        }
        const float rstd = rsqrtf(block_reduce<false>(ss, smem) * inv_n + eps);
        if (threadIdx.x == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
#pragma unroll
        for (int j = 0; j < kMaxVpt; ++j) {
            int v = threadIdx.x + j * kThreads;
            if (v < nvec) {
                float f[8], g[8], c[8];
                unpack8(xv[j], f);
                unpack8(__ldg(w + v), g);
                unpack8(__ldg(b + v), c);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = (f[i] - mean) * rstd * g[i] + c[i];
                st16(y + r * nvec + v, pack8(f));
            }
        }
    }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) with g = dy * w;  dw_partial[cta] = sum dy * xhat, db_partial[cta] = sum dy
__global__ void __launch_bounds__(kThreads) layernorm_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                                 const uint4* __restrict__ w, const float* __restrict__ mean_in,
                                                                 const float* __restrict__ rstd_in, uint4* __restrict__ dx,
                                                                 float* __restrict__ dw_partial, float* __restrict__ db_partial,
                                                                 long long rows, int nvec) {
    __shared__ float smem[32];
    // one 16-B vector per thread per pass keeps the accumulators in registers for rows up to kThreads * 8 * kMaxVpt columns
    float dw[kMaxVpt][8], db[kMaxVpt][8];
#pragma unroll
    for (int j = 0; j < kMaxVpt; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) { dw[j][i] = 0.f; db[j][i] = 0.f; }
    const float inv_n = 1.f / (float)(nvec * 8);
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const float mean = mean_in[r], rstd = rstd_in[r];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxVpt; ++j) {
            int v = threadIdx.x + j * kThreads;
            if (v < nvec) {
                float xh[8], g[8], wv[8];
                unpack8(ld16_stream(x + r * nvec + v), xh);
                unpack8(ld16_stream(dy + r * nvec + v), g);
                unpack8(__ldg(w + v), wv);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    xh[i] = (xh[i] - mean) * rstd;
                    dw[j][i] += g[i] * xh[i];
                    db[j][i] += g[i];
                    const float gw = g[i] * wv[i];
                    s1 += gw;
                    s2 += gw * xh[i];
                }
            }
        }
        s1 = block_reduce<false>(s1, smem) * inv_n;
        s2 = block_reduce<false>(s2, smem) * inv_n;
#pragma unroll
        for (int j = 0; j < kMaxVpt; ++j) {
            int v = threadIdx.x + j * kThreads;
            if (v < nvec) {       // re-read (L2-resident) instead of holding xhat and g*w for the whole row in registers
                float xh[8], g[8], wv[8], o[8];
                unpack8(ld16_stream(x + r * nvec + v), xh);
                unpack8(ld16_stream(dy + r * nvec + v), g);
                unpack8(__ldg(w + v), wv);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = rstd * (g[i] * wv[i] - s1 - (xh[i] - mean) * rstd * s2);
                st16(dx + r * nvec + v, pack8(o));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxVpt; ++j) {
        int v = threadIdx.x + j * kThreads;
        if (v < nvec) {
            float4* d = reinterpret_cast<float4*>(dw_partial + ((size_t)blockIdx.x * nvec + v) * 8);
            d[0] = make_float4(dw[j][0], dw[j][1], dw[j][2], dw[j][3]);
            d[1] = make_float4(dw[j][4], dw[j][5], dw[j][6], dw[j][7]);
            float4* e = reinterpret_cast<float4*>(db_partial + ((size_t)blockIdx.x * nvec + v) * 8);
            e[0] = make_float4(db[j][0], db[j][1], db[j][2], db[j][3]);
            e[1] = make_float4(db[j][4], db[j][5], db[j][6], db[j][7]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bias + GeLU (tanh form, Megatron's bias_gelu_impl: transformer.py:150-160 with bias_gelu_fusion; HF "gelu_new"), and the
// exact erf form (HF BERT "gelu"):  y = gelu(x + b);  backward: dx = dy * gelu'(x + b)   (dbias = column sums of dx)
// ---------------------------------------------------------------------------------------------
template <bool kTanh>
__device__ __forceinline__ float gelu_f(float v) {
    if (kTanh) return v * 0.5f * (1.f + tanhf(0.79788456f * v * (1.f + 0.044715f * v * v)));
    return v * 0.5f * (1.f + erff(v * 0.70710678f));
}
template <bool kTanh>
__device__ __forceinline__ float gelu_grad_f(float v) {
    if (kTanh) {
        const float t = tanhf(0.79788456f * v * (1.f + 0.044715f * v * v));
        return 0.5f * v * ((1.f - t * t) * (0.79788456f + 0.1070322243f * v * v)) + 0.5f * (1.f + t);
    }
    return 0.5f * (1.f + erff(v * 0.70710678f)) + v * 0.3989422804f * __expf(-0.5f * v * v);
}

template <bool kTanh, bool kBackward>
__global__ void __launch_bounds__(kThreads) bias_gelu_kernel(const uint4* __restrict__ x, const uint4* __restrict__ bias,
                                                             const uint4* __restrict__ dy, uint4* __restrict__ out,
                                                             long long rows, int cvec) {
    const size_t total = (size_t)rows * cvec, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t c = i % cvec;
        float f[8], b[8];
        unpack8(ld16_stream(x + i), f);
        if (bias != nullptr) {
            unpack8(__ldg(bias + c), b);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += b[k];
        }
        if (kBackward) {
            float d[8];
            unpack8(ld16_stream(dy + i), d);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = d[k] * gelu_grad_f<kTanh>(f[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = gelu_f<kTanh>(f[k]);
        }
        st16(out + i, pack8(f));
    }
}

// ---------------------------------------------------------------------------------------------
// swiglu: gate_up = [rows, 2*ffn] (gate = first half, up = second; transformer.py:122-124)
// ---------------------------------------------------------------------------------------------
// grid = (column blocks of a row, row groups): no index division; two rows per iteration = four 16-B loads in flight per thread
__global__ void __launch_bounds__(kThreads) swiglu_fwd_kernel(const uint4* __restrict__ gu, uint4* __restrict__ y,
                                                              long long rows, int fvec) {
    const int c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= fvec) return;
    const long long step = gridDim.y;
    for (long long r = blockIdx.y; r < rows; r += 2 * step) {
        const long long r2 = r + step;
        const bool two = r2 < rows;
        const uint4 g0 = ld16_stream(gu + r * 2 * fvec + c), u0 = ld16_stream(gu + r * 2 * fvec + fvec + c);
        uint4 g1 = g0, u1 = u0;
        if (two) { g1 = ld16_stream(gu + r2 * 2 * fvec + c); u1 = ld16_stream(gu + r2 * 2 * fvec + fvec + c); }
        float g[8], u[8];
        unpack8(g0, g); unpack8(u0, u);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = g[k] / (1.f + __expf(-g[k])) * u[k];
        st16(y + r * fvec + c, pack8(g));
        if (two) {
            unpack8(g1, g); unpack8(u1, u);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = g[k] / (1.f + __expf(-g[k])) * u[k];
            st16(y + r2 * fvec + c, pack8(g));
        }
    }
}

__global__ void __launch_bounds__(kThreads) swiglu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ gu,
                                                              uint4* __restrict__ dgu, long long rows, int fvec) {
    const int c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= fvec) return;
    const long long step = gridDim.y;
    for (long long r = blockIdx.y; r < rows; r += 2 * step) {
        const long long r2 = r + step;
        const bool two = r2 < rows;
        const uint4 g0 = ld16_stream(gu + r * 2 * fvec + c), u0 = ld16_stream(gu + r * 2 * fvec + fvec + c), d0 = ld16_stream(dy + r * fvec + c);
        uint4 g1 = g0, u1 = u0, d1 = d0;
        if (two) {
            g1 = ld16_stream(gu + r2 * 2 * fvec + c); u1 = ld16_stream(gu + r2 * 2 * fvec + fvec + c); d1 = ld16_stream(dy + r2 * fvec + c);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const long long rr = h ? r2 : r;
            float g[8], u[8], d[8], dg[8], du[8];
            unpack8(h ? g1 : g0, g); unpack8(h ? u1 : u0, u); unpack8(h ? d1 : d0, d);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float sg = 1.f / (1.f + __expf(-g[k]));
                du[k] = d[k] * g[k] * sg;
                dg[k] = d[k] * u[k] * sg * (1.f + g[k] * (1.f - sg));
            }
            st16(dgu + rr * 2 * fvec + c, pack8(dg));
            st16(dgu + rr * 2 * fvec + fvec + c, pack8(du));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fused QKV split + RoPE + [s,b,..] -> [b,s,heads,hn] relayout (and its exact transpose for backward).
// mixed: [s, b, ng, (r+2)*hn]  (per-group interleaved fused-QKV layout, transformer.py:733-756)
// q: [b, s, ng*r, hn]   k, v: [b, s, ng, hn]   cos/sin: [s, hn/2] fp32 (already offset for this rank)
// forward: q,k rotated by (+theta); backward: reads dq,dk,dv and writes dmixed rotated by (-theta).
// ---------------------------------------------------------------------------------------------
// One CTA walks tokens (s, b); a thread owns the same (group, head, half-head vector) of every token it visits, so the index
// decomposition is done once, outside the token loop, and all loads of a token are issued before the first use.
constexpr int kRopePairs = 2;   // (lo, hi) vector pairs per thread and token: covers ng * (r + 2) * hn / 16 <= 512 pairs (8192 columns)
__global__ void __launch_bounds__(kThreads) qkv_rope_kernel(uint4* __restrict__ mixed, uint4* __restrict__ q,
                                                            uint4* __restrict__ k, uint4* __restrict__ v,
                                                            const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                            long long s, long long b, int ng, int r, int hn, int backward) {
    const int half_vec = hn / 16;              // 16-B vectors in half a head
    const int heads = r + 2;
    const int pairs = ng * heads * half_vec;   // per token
    int pc[kRopePairs], pj[kRopePairs];
    size_t pm[kRopePairs], po[kRopePairs];     // offsets inside a token of the mixed row / of the destination row
    for (int e = 0; e < kRopePairs; ++e) {
        const int p = threadIdx.x + e * kThreads;
        const int c = p % half_vec, j = (p / half_vec) % heads, g = p / (half_vec * heads);
        pc[e] = c; pj[e] = p < pairs ? j : -1;
        pm[e] = ((size_t)g * heads + j) * (2 * half_vec) + c;
        po[e] = (j < r) ? ((size_t)g * r + j) * (2 * half_vec) + c : (size_t)g * (2 * half_vec) + c;
    }
    const size_t mixed_tok = (size_t)pairs * 2, q_tok = (size_t)ng * r * 2 * half_vec, kv_tok = (size_t)ng * 2 * half_vec;
    const long long tokens = s * b;
    for (int base = 0; base < pairs; base += kRopePairs * kThreads) {   // rows wider than kRopePairs * kThreads pairs: extra passes
        for (long long t = blockIdx.x; t < tokens; t += gridDim.x) {
            const long long si = t / b, bi = t - si * b;
            const size_t tok_out = (size_t)bi * s + si;                // q/k/v are [b, s, ...]
            uint4 lo4[kRopePairs], hi4[kRopePairs];
            uint4* dst[kRopePairs];
#pragma unroll
            for (int e = 0; e < kRopePairs; ++e) {
                if (pj[e] < 0) continue;
                uint4* m = mixed + (size_t)t * mixed_tok + pm[e];
                uint4* o = (pj[e] < r) ? q + tok_out * q_tok + po[e] : ((pj[e] == r) ? k : v) + tok_out * kv_tok + po[e];
                uint4* src = backward ? o : m;
                dst[e] = backward ? m : o;
                lo4[e] = ld16_stream(src);
                hi4[e] = ld16_stream(src + half_vec);
            }
#pragma unroll
            for (int e = 0; e < kRopePairs; ++e) {
                if (pj[e] < 0) continue;
                float lo[8], hi[8];
                unpack8(lo4[e], lo);
                unpack8(hi4[e], hi);
                if (pj[e] <= r) {  // q and k heads are rotated, v is copied
                    const float4* cp = reinterpret_cast<const float4*>(cos_t + si * (hn / 2) + pc[e] * 8);
                    const float4* sp = reinterpret_cast<const float4*>(sin_t + si * (hn / 2) + pc[e] * 8);
                    float4 c0 = __ldg(cp), c1 = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
                    const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float sg = backward ? -sn[i] : sn[i];
                        const float a = lo[i], bb = hi[i];
                        lo[i] = a * cs[i] - bb * sg;   // t*cos + rotate_half(t)*sin, rotate_half = (-x2, x1)
                        hi[i] = bb * cs[i] + a * sg;
                    }
                }
                st16(dst[e], pack8(lo));
                st16(dst[e] + half_vec, pack8(hi));
            }
        }
        if (base + kRopePairs * kThreads < pairs) {   // next pass: shift this thread's pairs
            for (int e = 0; e < kRopePairs; ++e) {
                const int p = threadIdx.x + e * kThreads + base + kRopePairs * kThreads;
                const int c = p % half_vec, j = (p / half_vec) % heads, g = p / (half_vec * heads);
                pc[e] = c; pj[e] = p < pairs ? j : -1;
                pm[e] = ((size_t)g * heads + j) * (2 * half_vec) + c;
                po[e] = (j < r) ? ((size_t)g * r + j) * (2 * half_vec) + c : (size_t)g * (2 * half_vec) + c;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// vocab-parallel cross-entropy (cross_entropy.py:14-152): three row kernels around two tiny all-reduces
// ---------------------------------------------------------------------------------------------
constexpr int kCeThreads = 512;

template <bool kBf16>
__device__ __forceinline__ int ce_load(const void* row, size_t v, float* f) {
    if (kBf16) { unpack8(ld16_stream(reinterpret_cast<const uint4*>(row) + v), f); return 8; }
    uint4 a = ld16_stream(reinterpret_cast<const uint4*>(row) + v);
    f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    return 4;
}

template <bool kBf16>
__global__ void __launch_bounds__(kCeThreads) ce_rowmax_kernel(const void* __restrict__ logits, float* __restrict__ rowmax,
                                                               long long rows, long long vocab) {
    __shared__ float smem[32];
    constexpr int E = kBf16 ? 8 : 4;
    const size_t nvec = vocab / E;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const char* row = reinterpret_cast<const char*>(logits) + (size_t)r * vocab * (kBf16 ? 2 : 4);
        float m = -INFINITY;
        for (size_t v = threadIdx.x; v < nvec; v += blockDim.x) {
            float f[8];
            ce_load<kBf16>(row, v, f);
#pragma unroll
            for (int i = 0; i < E; ++i) m = fmaxf(m, f[i]);
        }
        m = block_reduce<true>(m, smem);
        if (threadIdx.x == 0) rowmax[r] = m;
    }
}

// out2[r] = (sum_j exp(x_j - max_r), x_target - max_r if the target falls in [vocab_start, vocab_start+vocab) else 0)
template <bool kBf16>
__global__ void __launch_bounds__(kCeThreads) ce_sumexp_kernel(const void* __restrict__ logits, const long long* __restrict__ target,
                                                               const float* __restrict__ rowmax, float* __restrict__ out2,
                                                               long long rows, long long vocab, long long vocab_start) {
    __shared__ float smem[32];
    constexpr int E = kBf16 ? 8 : 4;
    const size_t nvec = vocab / E;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const char* row = reinterpret_cast<const char*>(logits) + (size_t)r * vocab * (kBf16 ? 2 : 4);
        const float m = rowmax[r];
        float sum = 0.f;
        for (size_t v = threadIdx.x; v < nvec; v += blockDim.x) {
            float f[8];
            ce_load<kBf16>(row, v, f);
#pragma unroll
            for (int i = 0; i < E; ++i) sum += __expf(f[i] - m);
        }
        sum = block_reduce<false>(sum, smem);
        if (threadIdx.x == 0) {
            const long long t = target[r] - vocab_start;
            float pred = 0.f;
            if (t >= 0 && t < vocab)
                pred = (kBf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(row)[t]) : reinterpret_cast<const float*>(row)[t]) - m;
            out2[2 * r] = sum;
            out2[2 * r + 1] = pred;
        }
    }
}

// in place: logits <- (softmax - onehot(target)) * grad_loss[r]
template <bool kBf16>
__global__ void __launch_bounds__(kCeThreads) ce_bwd_kernel(void* __restrict__ logits, const long long* __restrict__ target,
                                                            const float* __restrict__ rowmax, const float* __restrict__ sum2,
                                                            const float* __restrict__ grad_loss, long long rows, long long vocab,
                                                            long long vocab_start) {
    constexpr int E = kBf16 ? 8 : 4;
    const size_t nvec = vocab / E;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        char* row = reinterpret_cast<char*>(logits) + (size_t)r * vocab * (kBf16 ? 2 : 4);
        const float m = rowmax[r], inv = 1.f / sum2[2 * r], g = grad_loss[r];
        const long long t = target[r] - vocab_start;
        for (size_t v = threadIdx.x; v < nvec; v += blockDim.x) {
            float f[8];
            ce_load<kBf16>(row, v, f);
#pragma unroll
            for (int i = 0; i < E; ++i) {
                float p = __expf(f[i] - m) * inv;
                if ((long long)(v * E + i) == t) p -= 1.f;
                f[i] = p * g;
            }
            if (kBf16) st16(reinterpret_cast<uint4*>(row) + v, pack8(f));
            else reinterpret_cast<float4*>(row)[v] = make_float4(f[0], f[1], f[2], f[3]);
        }
    }
}

}  // namespace

#define BG_ALIGNED16(p) (((uintptr_t)(p) % 16) == 0)

extern "C" int bg_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t elems, float scale, int accumulate,
                       void* stream) {
    if (elems % 8) return fail(BG_EINVAL, "bg_cast: elems %zu must be a multiple of 8", elems);
    if (!BG_ALIGNED16(src) || !BG_ALIGNED16(dst)) return fail(BG_EINVAL, "bg_cast: pointers must be 16-B aligned");
    if (elems == 0) return BG_OK;
    const size_t nvec = elems / 8;
    int grid = local_grid(nvec, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
    const bool sb = src_dtype == BG_BF16, db = dst_dtype == BG_BF16;
    if (sb && db) cast_kernel<true, true><<<grid, kThreads, 0, st>>>(src, dst, nvec, scale, accumulate);
    else if (sb && !db) cast_kernel<true, false><<<grid, kThreads, 0, st>>>(src, dst, nvec, scale, accumulate);
    else if (!sb && db) cast_kernel<false, true><<<grid, kThreads, 0, st>>>(src, dst, nvec, scale, accumulate);
    else cast_kernel<false, false><<<grid, kThreads, 0, st>>>(src, dst, nvec, scale, accumulate);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

static int norm_args(long long rows, long long cols, const char* who) {
    if (rows < 0 || cols <= 0 || cols % 8) return fail(BG_EINVAL, "%s: cols %lld must be a positive multiple of 8", who, cols);
    if (cols / 8 > (long long)kMaxVpt * kThreads) return fail(BG_EUNSUPPORTED, "%s: cols %lld > %d", who, cols, kMaxVpt * kThreads * 8);
    return BG_OK;
}

extern "C" int bg_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, long long rows, long long cols, float eps,
                              void* stream) {
    int rc = norm_args(rows, cols, "bg_rmsnorm_fwd");
    if (rc) return rc;
    if (!BG_ALIGNED16(x) || !BG_ALIGNED16(w) || !BG_ALIGNED16(y)) return fail(BG_EINVAL, "bg_rmsnorm_fwd: 16-B alignment");
    if (rows == 0) return BG_OK;
    int grid = (int)(rows < g_tun.local_ctas ? rows : g_tun.local_ctas);
    const int nvec = (int)(cols / 8), vpt = (nvec + kThreads - 1) / kThreads;
#define BG_RMS_FWD(V) rmsnorm_fwd_kernel<V><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)w, (uint4*)y, rstd, rows, nvec, eps)
    if (vpt <= 1) BG_RMS_FWD(1); else if (vpt == 2) BG_RMS_FWD(2); else BG_RMS_FWD(4);
#undef BG_RMS_FWD
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw_partial,
                              long long rows, long long cols, int n_partial, void* stream) {
    int rc = norm_args(rows, cols, "bg_rmsnorm_bwd");
    if (rc) return rc;
    if (n_partial < 1) return fail(BG_EINVAL, "bg_rmsnorm_bwd: n_partial must be >= 1");
    if (!BG_ALIGNED16(dy) || !BG_ALIGNED16(x) || !BG_ALIGNED16(w) || !BG_ALIGNED16(dx) || !BG_ALIGNED16(dw_partial))
        return fail(BG_EINVAL, "bg_rmsnorm_bwd: 16-B alignment");
    const int nvec = (int)(cols / 8), vpt = (nvec + kThreads - 1) / kThreads;
#define BG_RMS_BWD(V) rmsnorm_bwd_kernel<V><<<n_partial, kThreads, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)x, (const uint4*)w, rstd, (uint4*)dx, dw_partial, rows, nvec)
    if (vpt <= 1) BG_RMS_BWD(1); else if (vpt == 2) BG_RMS_BWD(2); else BG_RMS_BWD(4);
#undef BG_RMS_BWD
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// (column blocks, row groups): about local_ctas CTAs in total, every CTA visiting >= 2 rows when there are enough of them
static dim3 swiglu_grid(long long rows, long long fvec) {
    const long long cb = (fvec + kThreads - 1) / kThreads;
    long long gy = g_tun.local_ctas / cb;
    if (gy < 1) gy = 1;
    if (gy > (rows + 1) / 2) gy = (rows + 1) / 2;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)cb, (unsigned)gy, 1);
}

extern "C" int bg_swiglu_fwd(const void* gate_up, void* y, long long rows, long long ffn, void* stream) {
    if (ffn <= 0 || ffn % 8) return fail(BG_EINVAL, "bg_swiglu_fwd: ffn %lld must be a multiple of 8", ffn);
    if (!BG_ALIGNED16(gate_up) || !BG_ALIGNED16(y)) return fail(BG_EINVAL, "bg_swiglu_fwd: 16-B alignment");
    if (rows == 0) return BG_OK;
    swiglu_fwd_kernel<<<swiglu_grid(rows, ffn / 8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)gate_up, (uint4*)y, rows, (int)(ffn / 8));
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_swiglu_bwd(const void* dy, const void* gate_up, void* dgate_up, long long rows, long long ffn, void* stream) {
    if (ffn <= 0 || ffn % 8) return fail(BG_EINVAL, "bg_swiglu_bwd: ffn %lld must be a multiple of 8", ffn);
    if (!BG_ALIGNED16(dy) || !BG_ALIGNED16(gate_up) || !BG_ALIGNED16(dgate_up)) return fail(BG_EINVAL, "bg_swiglu_bwd: 16-B alignment");
    if (rows == 0) return BG_OK;
    swiglu_bwd_kernel<<<swiglu_grid(rows, ffn / 8), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)gate_up, (uint4*)dgate_up, rows,
                                                                                       (int)(ffn / 8));
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_qkv_rope(void* mixed, void* q, void* k, void* v, const float* cos_t, const float* sin_t, long long s,
                           long long b, long long ng, long long r, long long hn, int backward, void* stream) {
    if (hn <= 0 || hn % 16) return fail(BG_EINVAL, "bg_qkv_rope: head dim %lld must be a multiple of 16", hn);
    if (ng < 1 || r < 1) return fail(BG_EINVAL, "bg_qkv_rope: bad head counts");
    if (!BG_ALIGNED16(mixed) || !BG_ALIGNED16(q) || !BG_ALIGNED16(k) || !BG_ALIGNED16(v) || !BG_ALIGNED16(cos_t) || !BG_ALIGNED16(sin_t))
        return fail(BG_EINVAL, "bg_qkv_rope: 16-B alignment");
    const size_t total = (size_t)s * b * ng * (r + 2) * (hn / 16);
    if (total == 0) return BG_OK;
    const long long tokens = s * b;
    qkv_rope_kernel<<<(int)(tokens < g_tun.local_ctas ? tokens : g_tun.local_ctas), kThreads, 0, (cudaStream_t)stream>>>(
        (uint4*)mixed, (uint4*)q, (uint4*)k, (uint4*)v, cos_t, sin_t, s, b, (int)ng, (int)r, (int)hn, backward);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

static int ce_args(int dtype, long long vocab, const void* logits, const char* who) {
    if (dtype != BG_BF16 && dtype != BG_F32) return fail(BG_EUNSUPPORTED, "%s: dtype %d", who, dtype);
    if (vocab <= 0 || vocab % 8) return fail(BG_EINVAL, "%s: local vocab %lld must be a multiple of 8", who, vocab);
    if (!BG_ALIGNED16(logits)) return fail(BG_EINVAL, "%s: 16-B alignment", who);
    return BG_OK;
}

extern "C" int bg_ce_rowmax(const void* logits, int dtype, float* rowmax, long long rows, long long vocab, void* stream) {
    int rc = ce_args(dtype, vocab, logits, "bg_ce_rowmax");
    if (rc) return rc;
    if (rows == 0) return BG_OK;
    int grid = (int)(rows < g_tun.local_ctas ? rows : g_tun.local_ctas);
    if (dtype == BG_BF16) ce_rowmax_kernel<true><<<grid, kCeThreads, 0, (cudaStream_t)stream>>>(logits, rowmax, rows, vocab);
    else ce_rowmax_kernel<false><<<grid, kCeThreads, 0, (cudaStream_t)stream>>>(logits, rowmax, rows, vocab);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_ce_sumexp(const void* logits, int dtype, const long long* target, const float* rowmax, float* out2,
                            long long rows, long long vocab, long long vocab_start, void* stream) {
    int rc = ce_args(dtype, vocab, logits, "bg_ce_sumexp");
    if (rc) return rc;
    if (rows == 0) return BG_OK;
    int grid = (int)(rows < g_tun.local_ctas ? rows : g_tun.local_ctas);
    if (dtype == BG_BF16) ce_sumexp_kernel<true><<<grid, kCeThreads, 0, (cudaStream_t)stream>>>(logits, target, rowmax, out2, rows, vocab, vocab_start);
    else ce_sumexp_kernel<false><<<grid, kCeThreads, 0, (cudaStream_t)stream>>>(logits, target, rowmax, out2, rows, vocab, vocab_start);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_ce_bwd(void* logits, int dtype, const long long* target, const float* rowmax, const float* sum2,
                         const float* grad_loss, long long rows, long long vocab, long long vocab_start, void* stream) {
    int rc = ce_args(dtype, vocab, logits, "bg_ce_bwd");
    if (rc) return rc;
    if (rows == 0) return BG_OK;
    int grid = (int)(rows < g_tun.local_ctas ? rows : g_tun.local_ctas);
    if (dtype == BG_BF16) ce_bwd_kernel<true><<<grid, kCeThreads, 0, (cudaStream_t)stream>>>(logits, target, rowmax, sum2, grad_loss, rows, vocab, vocab_start);
    else ce_bwd_kernel<false><<<grid, kCeThreads, 0, (cudaStream_t)stream>>>(logits, target, rowmax, sum2, grad_loss, rows, vocab, vocab_start);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, long long rows,
                                long long cols, float eps, void* stream) {
    if (cols % 8 || cols > (long long)kThreads * 8 * kMaxVpt) return fail(BG_EINVAL, "layernorm: cols %lld must be a multiple of 8 and <= %d", cols, kThreads * 8 * kMaxVpt);
    if (rows <= 0) return BG_OK;
    const int grid = (int)(rows < g_tun.local_ctas ? rows : g_tun.local_ctas);
    layernorm_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)w, (const uint4*)b, (uint4*)y, mean, rstd,
                                                                      rows, (int)(cols / 8), eps);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                                float* dw_partial, float* db_partial, long long rows, long long cols, int n_partial, void* stream) {
    if (cols % 8 || cols > (long long)kThreads * 8 * kMaxVpt) return fail(BG_EINVAL, "layernorm: cols %lld must be a multiple of 8 and <= %d", cols, kThreads * 8 * kMaxVpt);
    if (n_partial < 1) return fail(BG_EINVAL, "layernorm_bwd: n_partial must be >= 1");
    layernorm_bwd_kernel<<<n_partial, kThreads, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)x, (const uint4*)w, mean, rstd, (uint4*)dx,
                                                                           dw_partial, db_partial, rows, (int)(cols / 8));
    BG_CHECK_LAUNCH();
    return BG_OK;
}

extern "C" int bg_bias_gelu(const void* x, const void* bias, const void* dy, void* out, long long rows, long long cols, int tanh_form,
                            void* stream) {
    if (cols % 8) return fail(BG_EINVAL, "bias_gelu: cols %lld must be a multiple of 8", cols);
    if (rows <= 0) return BG_OK;
    const int grid = local_grid((size_t)rows * cols / 8, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
    const uint4 *xv = (const uint4*)x, *bv = (const uint4*)bias, *dv = (const uint4*)dy;
    if (dy == nullptr) {
        if (tanh_form) bias_gelu_kernel<true, false><<<grid, kThreads, 0, st>>>(xv, bv, dv, (uint4*)out, rows, (int)(cols / 8));
        else bias_gelu_kernel<false, false><<<grid, kThreads, 0, st>>>(xv, bv, dv, (uint4*)out, rows, (int)(cols / 8));
    } else {
        if (tanh_form) bias_gelu_kernel<true, true><<<grid, kThreads, 0, st>>>(xv, bv, dv, (uint4*)out, rows, (int)(cols / 8));
        else bias_gelu_kernel<false, true><<<grid, kThreads, 0, st>>>(xv, bv, dv, (uint4*)out, rows, (int)(cols / 8));
    }
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// loads every kernel of this file up front (see bg_preload_coll in bg_coll.cu)
int bg_preload_ops() {
#define K(f) reinterpret_cast<const void*>(&f)
    const void* kernels[] = {K((cast_kernel<true, true>)), K((cast_kernel<true, false>)), K((cast_kernel<false, true>)), K((cast_kernel<false, false>)),
                             K(rmsnorm_fwd_kernel<1>), K(rmsnorm_fwd_kernel<2>), K(rmsnorm_fwd_kernel<4>), K(rmsnorm_bwd_kernel<1>), K(rmsnorm_bwd_kernel<2>), K(rmsnorm_bwd_kernel<4>), K(layernorm_fwd_kernel), K(layernorm_bwd_kernel),
                             K((bias_gelu_kernel<true, false>)), K((bias_gelu_kernel<false, false>)), K((bias_gelu_kernel<true, true>)),
                             K((bias_gelu_kernel<false, true>)), K(swiglu_fwd_kernel), K(swiglu_bwd_kernel), K(qkv_rope_kernel),
                             K(ce_rowmax_kernel<true>), K(ce_rowmax_kernel<false>), K(ce_sumexp_kernel<true>), K(ce_sumexp_kernel<false>),
                             K(ce_bwd_kernel<true>), K(ce_bwd_kernel<false>)};
#undef K
    for (const void* k : kernels) {
        cudaFuncAttributes attr;
        BG_CUDA(cudaFuncGetAttributes(&attr, k));
    }
    return BG_OK;
}
