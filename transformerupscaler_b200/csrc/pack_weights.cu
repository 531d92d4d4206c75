// tu_pack_weights: state_dict (host fp32 tensors under the reference's names) -> packed device weights + TuModelWeights, in C++.
// This is the host-side repack a non-Python caller needs to produce what tu_forward consumes; it writes exactly
// the layouts transformerupscaler_b200/packing.py writes (tests/test_gpu_models.py::test_c_packer_matches_python_packer runs the same
// forward with both).  Reference parameter names and shapes: WindowTransformer/model.py:187-222, FastTransformer/model.py:189-229 +
// utils.py:43-98, ResidualTransformer/model.py:69-112.
//
// The caller owns everything: `device_buf` (tu_packed_weights_bytes() bytes of device memory) receives all packed tensors with ONE
// cudaMemcpyAsync on `stream` from a host staging buffer the function frees before returning (pageable source: the copy has left the
// host buffer when the call returns); `out` is a caller-allocated host struct whose TuModelWeights points into device_buf and into the
// struct itself (blocks[], host_finconv_wb) -- do not move it afterwards.
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "tu_common.cuh"

namespace tu {
namespace {

struct Lookup {
    const TuNamedTensor *t;
    int n;
    const TuNamedTensor *find(const std::string &name) const {
        for (int i = 0; i < n; ++i)
            if (t[i].name && name == t[i].name) return &t[i];
        return nullptr;
    }
};

static inline uint16_t f2bf(float f) {          // round to nearest even, like torch's .to(torch.bfloat16)
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);      // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// staging buffer: every packed tensor at a 256-byte aligned offset
struct Stage {
    std::vector<uint8_t> buf;
    char *dev;
    bool bf16_compute;
    size_t reserve(size_t bytes) {
        size_t off = (buf.size() + 255) & ~(size_t)255;
        buf.resize(off + bytes);
        return off;
    }
    const float *put_f32(const float *src, size_t n) {
        size_t off = reserve(n * 4);
        memcpy(buf.data() + off, src, n * 4);
        return (const float *)(dev + off);
    }
    const float *put_f32(const std::vector<float> &v) { return put_f32(v.data(), v.size()); }
    const void *put_bf16(const float *src, size_t n) {
        size_t off = reserve(n * 2);
        uint16_t *d = (uint16_t *)(buf.data() + off);
        for (size_t i = 0; i < n; ++i) d[i] = f2bf(src[i]);
        return dev + off;
    }
    // compute dtype T: bf16 or fp32
    const void *put_T(const float *src, size_t n) { return bf16_compute ? put_bf16(src, n) : (const void *)put_f32(src, n); }
    const void *put_T(const std::vector<float> &v) { return put_T(v.data(), v.size()); }
};

struct Packer {
    Lookup sd;
    Stage st;
    std::string err;
    const float *need(const std::string &name, size_t numel) {
        const TuNamedTensor *t = sd.find(name);
        if (!t || !t->data) { if (err.empty()) err = "missing tensor '" + name + "'"; return nullptr; }
        if ((size_t)t->numel != numel) {
            if (err.empty()) err = "tensor '" + name + "' has " + std::to_string(t->numel) + " elements, expected " + std::to_string(numel);
            return nullptr;
        }
        return (const float *)t->data;
    }
    // (co, 64, 3, 3) -> [tap][co][ci]
    const void *conv64(const std::string &name, int co) {
        const float *w = need(name, (size_t)co * 64 * 9);
        if (!w) return nullptr;
        std::vector<float> o((size_t)9 * co * 64);
        for (int c = 0; c < co; ++c)
            for (int ci = 0; ci < 64; ++ci)
                for (int t = 0; t < 9; ++t) o[((size_t)t * co + c) * 64 + ci] = w[((size_t)c * 64 + ci) * 9 + t];
        return st.put_T(o);
    }
    // (co, 3, 3, 3) -> (27, co): [(ky*3+kx)*3+ci][co]
    const float *conv_small_in(const std::string &name, int co) {
        const float *w = need(name, (size_t)co * 27);
        if (!w) return nullptr;
        std::vector<float> o((size_t)27 * co);
        for (int c = 0; c < co; ++c)
            for (int ci = 0; ci < 3; ++ci)
                for (int t = 0; t < 9; ++t) o[((size_t)t * 3 + ci) * co + c] = w[((size_t)c * 3 + ci) * 9 + t];
        return st.put_f32(o);
    }
    // (3, 64, 3, 3) -> (9, 64, 3)
    const float *conv_to3(const float *w) {
        std::vector<float> o(9 * 64 * 3);
        for (int c = 0; c < 3; ++c)
            for (int ci = 0; ci < 64; ++ci)
                for (int t = 0; t < 9; ++t) o[((size_t)t * 64 + ci) * 3 + c] = w[((size_t)c * 64 + ci) * 9 + t];
        return st.put_f32(o);
    }
    // (3, 64, 3, 3) -> bf16 (3 ky, 16 rows n = kx*4 + co [co < 3], 64 ci)
    const void *conv_to3_tc(const float *w) {
        std::vector<float> o(3 * 16 * 64, 0.f);
        for (int c = 0; c < 3; ++c)
            for (int ci = 0; ci < 64; ++ci)
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) o[((size_t)ky * 16 + kx * 4 + c) * 64 + ci] = w[((size_t)c * 64 + ci) * 9 + ky * 3 + kx];
        return st.put_bf16(o.data(), o.size());
    }
    // (3, 64, 3, 3) -> bf16 (3 kx, 3 blocks ky = 2..0, 16 rows co [co < 3], 64 ci)
    const void *conv_to3_stream(const float *w) {
        std::vector<float> o(3 * 3 * 16 * 64, 0.f);
        for (int c = 0; c < 3; ++c)
            for (int ci = 0; ci < 64; ++ci)
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) o[(((size_t)kx * 3 + (2 - ky)) * 16 + c) * 64 + ci] = w[((size_t)c * 64 + ci) * 9 + ky * 3 + kx];
        return st.put_bf16(o.data(), o.size());
    }
    const float *bias16(const float *b) {
        std::vector<float> o(16, 0.f);
        if (b) for (int i = 0; i < 3; ++i) o[i] = b[i];
        return st.put_f32(o);
    }
    // dense (heads, 64, 64) relative-position bias from the (225, heads) table (index = (yi - yj + 7) * 15 + (xi - xj + 7))
    std::vector<float> dense_rel(const float *table, const int64_t *index, int heads) {
        std::vector<float> o((size_t)heads * 4096);
        for (int i = 0; i < 64; ++i)
            for (int j = 0; j < 64; ++j) {
                int64_t idx = index ? index[i * 64 + j] : ((i / 8 - j / 8 + 7) * 15 + (i % 8 - j % 8 + 7));
                for (int h = 0; h < heads; ++h) o[((size_t)h * 64 + i) * 64 + j] = table[idx * heads + h];
            }
        return o;
    }
    // the same values in the mma.sync C-fragment order the fused stack kernels fetch (packing.py::frag_rel_bias):
    // [head][rg][n][lane = g*4 + tq][half][e] <- dense[head][rg*16 + half*8 + g][n*8 + tq*2 + e]
    static std::vector<float> frag_rel(const std::vector<float> &dense, int heads) {
        std::vector<float> o(dense.size());
        size_t k = 0;
        for (int h = 0; h < heads; ++h)
            for (int rg = 0; rg < 4; ++rg)
                for (int n = 0; n < 8; ++n)
                    for (int g = 0; g < 8; ++g)
                        for (int tq = 0; tq < 4; ++tq)
                            for (int half = 0; half < 2; ++half)
                                for (int e = 0; e < 2; ++e)
                                    o[k++] = dense[((size_t)h * 64 + rg * 16 + half * 8 + g) * 64 + n * 8 + tq * 2 + e];
        return o;
    }
};

// folded last up1 stage + up1_conv (packing.py::fold_up1), fp64.  Wf[vy][vx][(c r + i) r + j][ci][dy][dx], bf[vy][vx][o]
static void fold_up1(const float *w1f, const float *b1f, const float *w2f, int r, std::vector<double> &Wf, std::vector<double> &bf) {
    const int no = 3 * r * r;
    Wf.assign((size_t)9 * no * 64 * 25, 0.0);
    bf.assign((size_t)9 * no, 0.0);
    auto W1 = [&](int m, int ip, int jp, int ci, int ay, int ax) { return (double)w1f[(((size_t)(m * r * r + ip * r + jp) * 64 + ci) * 3 + ay) * 3 + ax]; };
    std::vector<double> contrib((size_t)3 * 64 * 9);
    for (int i = 0; i < r; ++i)
        for (int ky = 0; ky < 3; ++ky) {
            const int t = i + ky - 1;
            const int Dy = (t >= 0 ? t / r : -((-t + r - 1) / r)), ip = t - Dy * r;
            for (int j = 0; j < r; ++j)
                for (int kx = 0; kx < 3; ++kx) {
                    const int u = j + kx - 1;
                    const int Dx = (u >= 0 ? u / r : -((-u + r - 1) / r)), jp = u - Dx * r;
                    double cb[3] = {0, 0, 0};
                    std::fill(contrib.begin(), contrib.end(), 0.0);
                    for (int c = 0; c < 3; ++c)
                        for (int m = 0; m < 64; ++m) {
                            const double w2 = (double)w2f[((size_t)c * 64 + m) * 9 + ky * 3 + kx];
                            cb[c] += w2 * (double)b1f[m * r * r + ip * r + jp];
                            for (int ci = 0; ci < 64; ++ci)
                                for (int a = 0; a < 9; ++a) contrib[((size_t)c * 64 + ci) * 9 + a] += w2 * W1(m, ip, jp, ci, a / 3, a % 3);
                        }
                    for (int vy = 0; vy < 3; ++vy) {
                        if ((vy == 0 && ky == 0) || (vy == 2 && ky == 2)) continue;
                        for (int vx = 0; vx < 3; ++vx) {
                            if ((vx == 0 && kx == 0) || (vx == 2 && kx == 2)) continue;
                            for (int c = 0; c < 3; ++c) {
                                const int o = (c * r + i) * r + j;
                                bf[(size_t)(vy * 3 + vx) * no + o] += cb[c];
                                for (int ci = 0; ci < 64; ++ci)
                                    for (int ay = 0; ay < 3; ++ay)
                                        for (int ax = 0; ax < 3; ++ax)
                                            Wf[((((size_t)(vy * 3 + vx) * no + o) * 64 + ci) * 5 + (Dy + 1 + ay)) * 5 + (Dx + 1 + ax)] +=
                                                contrib[((size_t)c * 64 + ci) * 9 + ay * 3 + ax];
                            }
                        }
                    }
                }
        }
}

static int scale_of_slot(int s) { return s == 0 ? 2 : s == 1 ? 3 : s == 2 ? 4 : 6; }

}  // namespace
}  // namespace tu

using namespace tu;

extern "C" size_t tu_packed_weights_bytes(int model, int dim, int n_blocks, int compute_dtype) {
    if (model < 0 || model > 2 || (dim != 128 && dim != 192) || n_blocks < 1 || n_blocks > TU_MAX_BLOCKS) return 0;
    // generous closed-form bound: every tensor at most 4 bytes per element + padding, fused-stack copies and folded filters included
    size_t per_block = (size_t)12 * dim * dim * 4 * 2 + (size_t)(dim / 16) * 4096 * 4 * 2 + 64 * 1024;
    size_t fixed = (size_t)2 * dim * 4096 * 4 + 6 * 9 * 64 * 64 * 4 + (size_t)3600 * dim * 4 + (1 << 20);
    if (model == TU_MODEL_FAST) fixed += (size_t)(4 + 9 + 4 + 4 + 36) * 9 * 64 * 64 * 4 + (size_t)4 * (9 * 108 * 25 * 64 * 4 + 3 * 25 * 48 * 64 * 2) + (4 << 20);
    return fixed + per_block * n_blocks;
}

extern "C" int tu_pack_weights(int model, const TuNamedTensor *tensors, int n_tensors, int compute_dtype, void *device_buf, size_t device_bytes,
                               TuPackedModel *out, void *stream) {
    TU_CHECK_ARG(model >= 0 && model <= 2 && tensors && n_tensors > 0 && device_buf && out, "pack_weights: bad argument");
    TU_CHECK_ARG(compute_dtype == TU_F32 || compute_dtype == TU_BF16, "pack_weights: compute dtype must be TU_F32 or TU_BF16");
    Packer P;
    P.sd = Lookup{tensors, n_tensors};
    P.st.dev = (char *)device_buf;
    P.st.bf16_compute = compute_dtype == TU_BF16;
    const bool bf = P.st.bf16_compute, fast = model == TU_MODEL_FAST, resid = model == TU_MODEL_RESIDUAL;
    const TuNamedTensor *pe = P.sd.find("patch_embed.weight");
    TU_CHECK_ARG(pe && pe->numel % (64 * 64) == 0, "pack_weights: patch_embed.weight missing or malformed");
    const int dim = (int)(pe->numel / 4096), heads = dim / 16;
    TU_CHECK_ARG(dim == 128 || dim == 192, "pack_weights: transformer dim must be 128 or 192");
    const std::string bpre = resid ? "transformer_blocks." : "window_blocks.";
    int nb = 0;
    while (nb < TU_MAX_BLOCKS && P.sd.find(bpre + std::to_string(nb) + ".norm1.weight")) ++nb;
    TU_CHECK_ARG(nb > 0, "pack_weights: no transformer blocks in the state dict");
    memset(out, 0, sizeof(*out));
    TuModelWeights &mw = out->w;
    mw.model = model; mw.dim = dim; mw.heads = heads; mw.n_blocks = nb;

    mw.conv1_w = P.conv_small_in("conv1.weight", 64);
    { const float *b = P.need("conv1.bias", 64); if (b) mw.conv1_b = P.st.put_f32(b, 64); }
    if (bf) {
        const float *w = P.need("conv1.weight", 64 * 27);
        if (w) {
            std::vector<float> w64(64 * 64, 0.f);
            for (int co = 0; co < 64; ++co)
                for (int ci = 0; ci < 3; ++ci)
                    for (int t = 0; t < 9; ++t) w64[co * 64 + t * 3 + ci] = w[(co * 3 + ci) * 9 + t];
            mw.conv1_w64 = P.st.put_bf16(w64.data(), w64.size());
        }
    }
    mw.conv2_w = P.conv64("conv2.weight", 64);
    { const float *b = P.need("conv2.bias", 64); if (b) mw.conv2_b = P.st.put_f32(b, 64); }
    if (!fast) {
        mw.down_w = P.conv64("downsample.weight", 64);
        const float *b = P.need("downsample.bias", 64);
        if (b) mw.down_b = P.st.put_f32(b, 64);
    }
    {   // patch embed (dim, 64, 8, 8) -> (dim, ky, kx, ci)
        const float *w = P.need("patch_embed.weight", (size_t)dim * 4096);
        if (w) {
            std::vector<float> o((size_t)dim * 4096);
            for (int d = 0; d < dim; ++d)
                for (int ci = 0; ci < 64; ++ci)
                    for (int k = 0; k < 64; ++k) o[((size_t)d * 64 + k) * 64 + ci] = w[((size_t)d * 64 + ci) * 64 + k];
            mw.embed_w = P.st.put_T(o);
        }
        const float *b = P.need("patch_embed.bias", dim);
        if (b) mw.embed_b = P.st.put_f32(b, dim);
    }
    if (resid) {
        const float *pos = P.need("pos_embed", (size_t)3600 * dim);
        if (pos) mw.pos_embed = P.st.put_f32(pos, (size_t)3600 * dim);
    }
    {   // patch unembed: ConvTranspose2d weight (dim, 64, 8, 8) -> (ky, kx, co) x dim
        const float *w = P.need("patch_unembed.weight", (size_t)dim * 4096);
        if (w) {
            std::vector<float> o((size_t)4096 * dim);
            for (int d = 0; d < dim; ++d)
                for (int co = 0; co < 64; ++co)
                    for (int k = 0; k < 64; ++k) o[((size_t)k * 64 + co) * dim + d] = w[((size_t)d * 64 + co) * 64 + k];
            mw.unembed_w = P.st.put_T(o);
        }
        const float *b = P.need("patch_unembed.bias", 64);
        if (b) mw.unembed_b = P.st.put_f32(b, 64);
    }
    mw.dec1_w = P.conv64("decoder_conv1.weight", 64);
    { const float *b = P.need("decoder_conv1.bias", 64); if (b) mw.dec1_b = P.st.put_f32(b, 64); }
    {
        const float *w = P.need("decoder_conv2.weight", 3 * 64 * 9), *b = P.need("decoder_conv2.bias", 3);
        if (w && b) {
            mw.dec2_w = P.conv_to3(w);
            mw.dec2_b = P.st.put_f32(b, 3);
            if (bf) {
                mw.dec2_w16 = P.conv_to3_tc(w);
                mw.dec2_wst = P.conv_to3_stream(w);
                mw.dec2_b16 = P.bias16(b);
            }
        }
    }

    // ---- transformer blocks (q rows and q bias pre-scaled by head_dim^-0.5 = 0.25: exact in fp32 and bf16)
    std::vector<std::vector<float>> qw_s(nb), qb_s(nb), rel_s(nb);
    for (int i = 0; i < nb && P.err.empty(); ++i) {
        const std::string p = bpre + std::to_string(i) + ".";
        TuBlockWeights &bw = out->blocks[i];
        const float *t;
        if ((t = P.need(p + "norm1.weight", dim))) bw.ln1_w = P.st.put_f32(t, dim);
        if ((t = P.need(p + "norm1.bias", dim))) bw.ln1_b = P.st.put_f32(t, dim);
        if ((t = P.need(p + "norm2.weight", dim))) bw.ln2_w = P.st.put_f32(t, dim);
        if ((t = P.need(p + "norm2.bias", dim))) bw.ln2_b = P.st.put_f32(t, dim);
        const float *qw = P.need(p + (resid ? "attn.in_proj_weight" : "attn.qkv.weight"), (size_t)3 * dim * dim);
        const float *qb = P.need(p + (resid ? "attn.in_proj_bias" : "attn.qkv.bias"), (size_t)3 * dim);
        const float *pw = P.need(p + (resid ? "attn.out_proj.weight" : "attn.proj.weight"), (size_t)dim * dim);
        const float *pb = P.need(p + (resid ? "attn.out_proj.bias" : "attn.proj.bias"), dim);
        const float *w1 = P.need(p + "mlp.0.weight", (size_t)4 * dim * dim), *b1 = P.need(p + "mlp.0.bias", (size_t)4 * dim);
        const float *w2 = P.need(p + "mlp.2.weight", (size_t)4 * dim * dim), *b2 = P.need(p + "mlp.2.bias", dim);
        if (!P.err.empty()) break;
        qw_s[i].assign(qw, qw + (size_t)3 * dim * dim);
        qb_s[i].assign(qb, qb + (size_t)3 * dim);
        for (size_t k = 0; k < (size_t)dim * dim; ++k) qw_s[i][k] *= 0.25f;
        for (int k = 0; k < dim; ++k) qb_s[i][k] *= 0.25f;
        if (!resid) {
            const float *table = P.need(p + "attn.relative_position_bias_table", (size_t)225 * heads);
            const TuNamedTensor *ix = P.sd.find(p + "attn.relative_position_index");
            if (!table) break;
            rel_s[i] = P.dense_rel(table, ix ? (const int64_t *)ix->data : nullptr, heads);
            bw.rel_bias = P.st.put_f32(rel_s[i]);
        }
        bw.qkv_w = P.st.put_T(qw_s[i]); bw.qkv_b = P.st.put_f32(qb_s[i]);
        bw.proj_w = P.st.put_T(pw, (size_t)dim * dim); bw.proj_b = P.st.put_f32(pb, dim);
        bw.fc1_w = P.st.put_T(w1, (size_t)4 * dim * dim); bw.fc1_b = P.st.put_f32(b1, (size_t)4 * dim);
        bw.fc2_w = P.st.put_T(w2, (size_t)4 * dim * dim); bw.fc2_b = P.st.put_f32(b2, dim);
    }
    mw.blocks = out->blocks;

    // ---- fused window stack (bf16, WindowTransformer / FastTransformer): weight slabs in MMA consumption order
    if (bf && !resid && P.err.empty()) {
        std::vector<float> slabs, pars, rels;
        std::vector<double> c(dim, 0.0);
        auto slab = [&](const float *w, int ld, int r0, int nrows, int k0) {      // rows r0.. x columns k0..k0+64 of a row-major matrix
            for (int r = 0; r < nrows; ++r)
                for (int k = 0; k < 64; ++k) slabs.push_back(w[(size_t)(r0 + r) * ld + k0 + k]);
        };
        for (int i = 0; i < nb; ++i) {
            const std::string p = bpre + std::to_string(i) + ".";
            const float *qw = qw_s[i].data(), *qb = qb_s[i].data();
            const float *pw = P.need(p + "attn.proj.weight", (size_t)dim * dim), *pb = P.need(p + "attn.proj.bias", dim);
            const float *w1 = P.need(p + "mlp.0.weight", (size_t)4 * dim * dim), *b1 = P.need(p + "mlp.0.bias", (size_t)4 * dim);
            const float *w2 = P.need(p + "mlp.2.weight", (size_t)4 * dim * dim), *b2 = P.need(p + "mlp.2.bias", dim);
            std::vector<float> qb_order;
            if (dim == 128) {
                for (int nc = 0; nc < 3; ++nc)
                    for (int ks = 0; ks < 2; ++ks) slab(qw, dim, nc * 128, 128, ks * 64);
                for (int ks = 0; ks < 2; ++ks) slab(pw, dim, 0, 128, ks * 64);
                for (int h = 0; h < 2; ++h) {
                    for (int nc = 0; nc < 2; ++nc)
                        for (int ks = 0; ks < 2; ++ks) slab(w1, dim, h * 256 + nc * 128, 128, ks * 64);
                    for (int ks = 0; ks < 4; ++ks) slab(w2, 4 * dim, 0, 128, h * 256 + ks * 64);
                }
                qb_order.assign(qb, qb + 3 * dim);
            } else {
                // dim 192: per group g of two heads and K-slab: rows q | k | v (32 each); then proj (3 slabs of 192 rows)
                for (int g = 0; g < 6; ++g) {
                    for (int ks = 0; ks < 3; ++ks)
                        for (int s = 0; s < 3; ++s) slab(qw, dim, s * dim + g * 32, 32, ks * 64);
                    for (int s = 0; s < 3; ++s)
                        for (int r = 0; r < 32; ++r) qb_order.push_back(qb[s * dim + g * 32 + r]);
                }
                for (int ks = 0; ks < 3; ++ks) slab(pw, dim, 0, 192, ks * 64);
                // MLP in six chunks of 128 hidden units: fc1 chunk = 3 slabs [128 n x 64 k], fc2 chunk = 2 slabs [192 n x 64 k];
                // order fc1 c0, fc1 c1, then per chunk c: fc2 c, fc1 c+2 (packing.py::_pack_fused_stack192)
                auto fc1_chunk = [&](int cc) { for (int ks = 0; ks < 3; ++ks) slab(w1, dim, cc * 128, 128, ks * 64); };
                auto fc2_chunk = [&](int cc) { for (int ks = 0; ks < 2; ++ks) slab(w2, 4 * dim, 0, 192, cc * 128 + ks * 64); };
                fc1_chunk(0); fc1_chunk(1);
                for (int cc = 0; cc < 6; ++cc) {
                    fc2_chunk(cc);
                    if (cc + 2 < 6) fc1_chunk(cc + 2);
                }
            }
            std::vector<double> c0 = c, c1(dim);
            for (int k = 0; k < dim; ++k) { c1[k] = c0[k] + (double)pb[k]; c[k] = c1[k] + (double)b2[k]; }
            const float *n1w = P.need(p + "norm1.weight", dim), *n1b = P.need(p + "norm1.bias", dim);
            const float *n2w = P.need(p + "norm2.weight", dim), *n2b = P.need(p + "norm2.bias", dim);
            for (int k = 0; k < dim; ++k) pars.push_back((float)c0[k]);
            pars.insert(pars.end(), n1w, n1w + dim); pars.insert(pars.end(), n1b, n1b + dim);
            pars.insert(pars.end(), qb_order.begin(), qb_order.end());
            for (int k = 0; k < dim; ++k) pars.push_back((float)c1[k]);
            pars.insert(pars.end(), n2w, n2w + dim); pars.insert(pars.end(), n2b, n2b + dim);
            pars.insert(pars.end(), b1, b1 + 4 * dim);
            { const std::vector<float> fr = Packer::frag_rel(rel_s[i], heads); rels.insert(rels.end(), fr.begin(), fr.end()); }
        }
        for (int k = 0; k < dim; ++k) pars.push_back((float)c[k]);
        mw.stack_w = P.st.put_bf16(slabs.data(), slabs.size());
        mw.stack_p = P.st.put_f32(pars);
        mw.stack_rel = P.st.put_f32(rels);
    }

    // ---- ResidualTransformer (bf16, dim 128): the 24 slabs per layer of tc/residual_block_tcgen05.cu and 1792 parameters per layer
    if (bf && resid && dim == 128 && P.err.empty()) {
        std::vector<float> slabs, pars;
        auto slab = [&](const float *w, int ld, int r0, int nrows, int k0) {
            for (int r = 0; r < nrows; ++r)
                for (int k = 0; k < 64; ++k) slabs.push_back(w[(size_t)(r0 + r) * ld + k0 + k]);
        };
        for (int i = 0; i < nb; ++i) {
            const std::string p = bpre + std::to_string(i) + ".";
            const float *qw = qw_s[i].data(), *qb = qb_s[i].data();
            const float *pw = P.need(p + "attn.out_proj.weight", (size_t)dim * dim), *pb = P.need(p + "attn.out_proj.bias", dim);
            const float *w1 = P.need(p + "mlp.0.weight", (size_t)4 * dim * dim), *b1 = P.need(p + "mlp.0.bias", (size_t)4 * dim);
            const float *w2 = P.need(p + "mlp.2.weight", (size_t)4 * dim * dim), *b2 = P.need(p + "mlp.2.bias", dim);
            const float *n1w = P.need(p + "norm1.weight", dim), *n1b = P.need(p + "norm1.bias", dim);
            const float *n2w = P.need(p + "norm2.weight", dim), *n2b = P.need(p + "norm2.bias", dim);
            for (int nc = 0; nc < 3; ++nc)
                for (int ks = 0; ks < 2; ++ks) slab(qw, dim, nc * 128, 128, ks * 64);
            for (int ks = 0; ks < 2; ++ks) slab(pw, dim, 0, 128, ks * 64);
            for (int h = 0; h < 2; ++h) {
                for (int nc = 0; nc < 2; ++nc)
                    for (int ks = 0; ks < 2; ++ks) slab(w1, dim, h * 256 + nc * 128, 128, ks * 64);
                for (int ks = 0; ks < 4; ++ks) slab(w2, 4 * dim, 0, 128, h * 256 + ks * 64);
            }
            pars.insert(pars.end(), dim, 0.f);
            pars.insert(pars.end(), n1w, n1w + dim); pars.insert(pars.end(), n1b, n1b + dim);
            pars.insert(pars.end(), qb, qb + 3 * dim);
            pars.insert(pars.end(), pb, pb + dim);
            pars.insert(pars.end(), n2w, n2w + dim); pars.insert(pars.end(), n2b, n2b + dim);
            pars.insert(pars.end(), b1, b1 + 4 * dim);
            for (int k = 0; k < dim; ++k) pars.push_back((float)((double)pb[k] + (double)b2[k]));
        }
        mw.stack_w = P.st.put_bf16(slabs.data(), slabs.size());
        mw.stack_p = P.st.put_f32(pars);
    }

    // ---- FastTransformer: sub-pixel branches, folded up1 stage, 3 -> 3 tail
    if (fast && P.err.empty()) {
        const float *w2c = P.need("up1_conv.conv.weight", 3 * 64 * 9);
        for (int slot = 0; slot < 4 && P.err.empty(); ++slot) {
            const int s = scale_of_slot(slot);
            const int nst = s == 4 ? 2 : 1;
            for (int si = 0; si < nst; ++si) {
                const int idx = s == 4 ? si * 2 : 0, r = s == 4 ? 2 : s;
                const std::string pu = "up1.upsamplers." + std::to_string(s) + "." + std::to_string(idx);
                const std::string pf = "final_upscale.upsamplers." + std::to_string(s) + "." + std::to_string(idx);
                const float *w = P.need(pu + ".weight", (size_t)64 * r * r * 64 * 9), *b = P.need(pu + ".bias", (size_t)64 * r * r);
                if (!w || !b) break;
                // (64 r^2, 64, 3, 3), out channel o = c*r^2 + phase -> [phase][tap][c][ci]
                std::vector<float> wp((size_t)r * r * 9 * 64 * 64), bp((size_t)64 * r * r);
                for (int c = 0; c < 64; ++c)
                    for (int ph = 0; ph < r * r; ++ph) {
                        bp[(size_t)ph * 64 + c] = b[c * r * r + ph];
                        for (int ci = 0; ci < 64; ++ci)
                            for (int t = 0; t < 9; ++t)
                                wp[(((size_t)ph * 9 + t) * 64 + c) * 64 + ci] = w[(((size_t)(c * r * r + ph)) * 64 + ci) * 9 + t];
                    }
                mw.up1[slot][si].w = P.st.put_T(wp);
                mw.up1[slot][si].b = P.st.put_f32(bp);
                mw.up1[slot][si].r = r;
                mw.fin[slot][si].w = P.conv_small_in(pf + ".weight", 3 * r * r);
                const float *fb = P.need(pf + ".bias", (size_t)3 * r * r);
                if (fb) mw.fin[slot][si].b = P.st.put_f32(fb, (size_t)3 * r * r);
                mw.fin[slot][si].r = r;
                if (bf && si == nst - 1 && w2c) {      // folded last up1 stage + up1_conv (tensor-core path only)
                    std::vector<double> Wf, bfv;
                    fold_up1(w, b, w2c, r, Wf, bfv);
                    const int no3 = 3 * r * r;
                    const int NO = r == 2 ? 16 : r == 3 ? 32 : 48, rpc = r == 2 ? 6 : r == 3 ? 9 : 6, nchunk = r == 6 ? 3 : 1;
                    // interior filter (vy = vx = 1) -> bank (nchunk, 5 kx, 5 blocks ky = 4..0, NO, 64) and padded bias (nchunk * NO)
                    std::vector<float> bank((size_t)nchunk * 25 * NO * 64, 0.f), bias((size_t)nchunk * NO, 0.f);
                    for (int ch = 0; ch < nchunk; ++ch)
                        for (int qrow = 0; qrow < rpc; ++qrow)
                            for (int j = 0; j < r; ++j) {
                                const int o = (ch * rpc + qrow) * r + j, n = qrow * r + j;
                                bias[(size_t)ch * NO + n] = (float)bfv[(size_t)4 * no3 + o];
                                for (int kx = 0; kx < 5; ++kx)
                                    for (int blk = 0; blk < 5; ++blk)
                                        for (int ci = 0; ci < 64; ++ci)
                                            bank[((((size_t)ch * 5 + kx) * 5 + blk) * NO + n) * 64 + ci] =
                                                (float)Wf[((((size_t)4 * no3 + o) * 64 + ci) * 5 + (4 - blk)) * 5 + kx];
                            }
                    // ring filters: (9, 3r^2, 25 taps dy*5+dx, 64 ci)
                    std::vector<float> ring((size_t)9 * no3 * 25 * 64), ringb((size_t)9 * no3);
                    for (int v = 0; v < 9; ++v)
                        for (int o = 0; o < no3; ++o) {
                            ringb[(size_t)v * no3 + o] = (float)bfv[(size_t)v * no3 + o];
                            for (int ci = 0; ci < 64; ++ci)
                                for (int tp = 0; tp < 25; ++tp)
                                    ring[(((size_t)v * no3 + o) * 25 + tp) * 64 + ci] = (float)Wf[(((size_t)v * no3 + o) * 64 + ci) * 25 + tp];
                        }
                    TuUpFold &uf = mw.upfold[slot];
                    uf.w = P.st.put_bf16(bank.data(), bank.size());
                    uf.b = P.st.put_f32(bias);
                    uf.ring_w = P.st.put_f32(ring);
                    uf.ring_b = P.st.put_f32(ringb);
                    uf.r = r;
                }
            }
        }
        if (w2c) {
            mw.up1conv_w = P.conv_to3(w2c);
            if (bf) {
                mw.up1conv_w16 = P.conv_to3_tc(w2c);
                mw.up1conv_wst = P.conv_to3_stream(w2c);
                mw.up1conv_b16 = P.bias16(nullptr);
            }
        }
        mw.finconv_w = P.conv_small_in("final_upscale_conv.weight", 3);
        const float *fw = P.need("final_upscale_conv.weight", 81), *fb = P.need("final_upscale_conv.bias", 3);
        if (fw && fb) {
            mw.finconv_b = P.st.put_f32(fb, 3);
            for (int c = 0; c < 3; ++c)
                for (int ci = 0; ci < 3; ++ci)
                    for (int t = 0; t < 9; ++t) out->host_finconv_wb[(t * 3 + ci) * 3 + c] = fw[(c * 3 + ci) * 9 + t];
            for (int c = 0; c < 3; ++c) out->host_finconv_wb[81 + c] = fb[c];
            mw.host_finconv_wb = out->host_finconv_wb;
        }
    }
    if (!P.err.empty()) {
        set_error("tu: pack_weights: " + P.err);
        return TU_ERR_ARG;
    }
    if (P.st.buf.size() > device_bytes) {
        set_error("tu: pack_weights: device buffer too small: need " + std::to_string(P.st.buf.size()) + " bytes");
        return TU_ERR_WORKSPACE;
    }
    out->device_bytes_used = P.st.buf.size();
    cudaError_t e = cudaMemcpyAsync(device_buf, P.st.buf.data(), P.st.buf.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "pack_weights upload");
    return TU_OK;
}
