"""state_dict -> packed device weights for libtu_b200 (layouts documented in include/tu_b200.h).

One-off host-side plumbing (torch permutes / casts on the model's device), cached per
(parameter versions, compute dtype).  ``T`` below is the compute dtype (fp32 or bf16); biases,
LayerNorm affine, the dense relative-position bias and pos_embed stay fp32.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch

from . import _lib

SCALES = (2, 3, 4, 6)


def dense_rel_bias_t(table: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """(heads, query i, key j) fp32 from the (225, heads) table and the (64,64) index buffer
    (reference: WindowTransformer/model.py:118-122 gathers table[index] -> (i, j, heads) -> permute(2,0,1))."""
    b = table.float()[index.reshape(-1).long()].reshape(64, 64, -1)      # (i, j, h)
    return b.permute(2, 0, 1).contiguous()                               # (h, i, j)


def frag_rel_bias(dense: torch.Tensor) -> torch.Tensor:
    """Dense (heads, 64, 64) bias in the order the fused stack kernels read it (window_stack{,192}_tcgen05.cu): per (head, 16-row
    group rg, key octet n) the 32 lanes of a warp each own four floats -- (row rg*16 + g, columns n*8 + tq*2, +1) and the same
    columns of row rg*16 + g + 8, lane = g*4 + tq: the C fragment of mma.sync m16n8k16 -- so a fetch is one coalesced 16-byte
    load per lane instead of sixteen 8-byte loads scattered over eight cache lines each."""
    h = dense.shape[0]
    v = dense.reshape(h, 4, 2, 8, 8, 4, 2)                 # (h, rg, half, g, n, tq, e)
    return v.permute(0, 1, 4, 3, 5, 2, 6).contiguous().reshape(h, 64, 64)   # (h, rg, n, g, tq, half, e)


FOLD_CFG = {2: (16, 6, 1), 3: (32, 9, 1), 6: (48, 6, 3)}      # r -> (NO, (c,i) rows per chunk, chunks); tc/upfold_stream_tcgen05.cu


def fold_up1(w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, r: int):
    """Compose the last up1 stage with up1_conv (FastTransformer/utils.py:43-98, 13-40; model.py:264-265).

    w1 (64 r^2, 64, 3, 3), b1 (64 r^2): Conv2d before nn.PixelShuffle(r); w2 (3, 64, 3, 3): up1_conv (no bias).
    High-res pixel (y r + i, x r + j), channel c:
        out = sum_{m,ky,kx} w2[c,m,ky,kx] * U[m, y r + i + ky - 1, x r + j + kx - 1],   U = PixelShuffle(conv(F; w1, b1)),
    and U[m, y' r + i', x' r + j'] = b1[m r^2 + i' r + j'] + sum_{ci,ay,ax} w1[m r^2 + i' r + j', ci, ay, ax] F[ci, y'+ay-1, x'+ax-1],
    so out is a 5x5 convolution of F.  Returns (Wf, bf) in fp64 for the nine border cases (vy, vx) in {first, interior, last}^2
    of the HIGH-RES row / column: the reference zero-pads U, so on the first (last) row the ky = 0 (2) taps are absent.
    Wf: (3, 3, 3 r^2, 64, 5, 5) indexed [vy, vx, (c r + i) r + j, ci, dy + 2, dx + 2]; bf: (3, 3, 3 r^2).
    """
    w1 = w1.detach().double().cpu().reshape(64, r, r, 64, 3, 3)       # (m, i', j', ci, ay, ax)
    b1 = b1.detach().double().cpu().reshape(64, r, r)
    w2 = w2.detach().double().cpu()                                   # (c, m, ky, kx)
    Wf = torch.zeros(3, 3, 3, r, r, 64, 5, 5, dtype=torch.float64)     # (vy, vx, c, i, j, ci, dy, dx)
    bf = torch.zeros(3, 3, 3, r, r, dtype=torch.float64)
    for i in range(r):
        for ky in range(3):
            t = i + ky - 1
            Dy, ip = t // r, t % r
            for j in range(r):
                for kx in range(3):
                    u = j + kx - 1
                    Dx, jp = u // r, u % r
                    contrib = torch.einsum("cm,mkab->ckab", w2[:, :, ky, kx], w1[:, ip, jp])      # (c, ci, ay, ax)
                    cb = w2[:, :, ky, kx] @ b1[:, ip, jp]                                          # (c)
                    for vy in range(3):
                        if (vy == 0 and ky == 0) or (vy == 2 and ky == 2):
                            continue
                        for vx in range(3):
                            if (vx == 0 and kx == 0) or (vx == 2 and kx == 2):
                                continue
                            Wf[vy, vx, :, i, j, :, Dy + 1:Dy + 4, Dx + 1:Dx + 4] += contrib
                            bf[vy, vx, :, i, j] += cb
    return Wf.reshape(3, 3, 3 * r * r, 64, 5, 5), bf.reshape(3, 3, 3 * r * r)


def pack_fold_bank(Wf: torch.Tensor, bf: torch.Tensor, r: int):
    """Interior folded filter -> the tensor-core bank (nchunk, 5 kx, 5 blocks ky = 4..0, NO, 64) and the padded bias (nchunk*NO)."""
    NO, rpc, nchunk = FOLD_CFG[r]
    W = Wf[1, 1].reshape(3 * r, r, 64, 5, 5)              # ((c, i) row, j, ci, dy, dx)
    B = bf[1, 1].reshape(3 * r, r)
    bank = torch.zeros(nchunk, 5, 5, NO, 64, dtype=torch.float64)
    bias = torch.zeros(nchunk, NO, dtype=torch.float64)
    for ch in range(nchunk):
        rows = W[ch * rpc:(ch + 1) * rpc]                 # (q, j, ci, dy, dx)
        n = rows.shape[0] * r
        # bank[ch, kx, blk, q*r + j, ci] = W[q, j, ci, dy = 4 - blk, dx = kx]
        bank[ch, :, :, :n] = rows.flip(3).permute(4, 3, 0, 1, 2).reshape(5, 5, n, 64)
        bias[ch, :n] = B[ch * rpc:(ch + 1) * rpc].reshape(-1)
    return bank, bias.reshape(-1)


class PackedWeights:
    """Keeps the packed tensors alive and exposes the TuModelWeights struct."""

    def __init__(self, model: str, sd: Dict[str, torch.Tensor], dtype: torch.dtype, device: torch.device):
        self.model, self.dtype, self.device = model, dtype, device
        self.keep: List[torch.Tensor] = []
        f32 = torch.float32

        def dev(t: torch.Tensor, dt) -> torch.Tensor:
            t = t.detach().to(device=device, dtype=dt).contiguous()
            self.keep.append(t)
            return t

        def ptr(t) -> int:
            return 0 if t is None else t.data_ptr()

        def conv64(w):      # (64*nchunk? , 64, 3, 3) plain conv -> [tap][co][ci]
            return dev(w.float().permute(2, 3, 0, 1).reshape(9, w.shape[0], 64), dtype)

        def conv_small_in(w):   # (co, 3, 3, 3) -> (27, co): [(ky*3+kx)*3+ci][co]
            return dev(w.float().permute(2, 3, 1, 0).reshape(27, w.shape[0]), f32)

        def conv_to3(w):        # (3, 64, 3, 3) -> (9, 64, 3)
            return dev(w.float().permute(2, 3, 1, 0).reshape(9, 64, 3), f32)

        def conv_to3_tc(w):     # (3, 64, 3, 3) -> bf16 (3 ky, 16 rows n = kx*4 + co [co < 3], 64 ci) for the tensor-core head
            t = torch.zeros(3, 4, 4, 64, dtype=torch.float32, device=w.device)       # (ky, kx, co, ci)
            t[:, :3, :3] = w.float().permute(2, 3, 0, 1)                             # (ky, kx, co, ci)
            return dev(t.reshape(3, 16, 64), torch.bfloat16)

        def conv_to3_stream(w):  # (3, 64, 3, 3) -> bf16 (3 kx, 3 blocks ky = 2..0, 16 rows co [co < 3], 64 ci): streaming head
            t = torch.zeros(3, 3, 16, 64, dtype=torch.float32, device=w.device)
            t[:, :, :3] = w.float().permute(3, 2, 0, 1).flip(1)                      # (kx, ky, co, ci) with ky reversed
            return dev(t, torch.bfloat16)

        def bias16(b):
            t = torch.zeros(16, dtype=torch.float32)
            if b is not None:
                t[:3] = b.detach().float().cpu()
            return dev(t, f32)

        mw = _lib.TuModelWeights()
        mw.model = _lib.MODEL_IDS[model]
        fast, resid = model == "FastTransformer", model == "ResidualTransformer"
        dim = sd["patch_embed.weight"].shape[0]
        heads = dim // 16
        bprefix = "transformer_blocks." if resid else "window_blocks."
        nb = 1 + max(int(k[len(bprefix):].split(".")[0]) for k in sd if k.startswith(bprefix))
        mw.dim, mw.heads, mw.n_blocks = dim, heads, nb

        mw.conv1_w = ptr(conv_small_in(sd["conv1.weight"]))
        mw.conv1_b = ptr(dev(sd["conv1.bias"], f32))
        if dtype == torch.bfloat16:     # tensor-core stem: (64 co, 64 k), k = (ky*3+kx)*3+ci, zero-padded from 27
            w64 = torch.zeros(64, 64, dtype=torch.float32, device=sd["conv1.weight"].device)
            w64[:, :27] = sd["conv1.weight"].float().permute(0, 2, 3, 1).reshape(64, 27)
            mw.conv1_w64 = ptr(dev(w64, torch.bfloat16))
        mw.conv2_w = ptr(conv64(sd["conv2.weight"]))
        mw.conv2_b = ptr(dev(sd["conv2.bias"], f32))
        if not fast:
            mw.down_w = ptr(conv64(sd["downsample.weight"]))
            mw.down_b = ptr(dev(sd["downsample.bias"], f32))
        # patch embed (dim, 64, 8, 8) -> (dim, ky, kx, ci)
        mw.embed_w = ptr(dev(sd["patch_embed.weight"].float().permute(0, 2, 3, 1).reshape(dim, 4096), dtype))
        mw.embed_b = ptr(dev(sd["patch_embed.bias"], f32))
        if resid:
            mw.pos_embed = ptr(dev(sd["pos_embed"].reshape(-1, dim), f32))
        # patch unembed: ConvTranspose2d weight (dim, 64, 8, 8) -> (ky, kx, co) x dim
        mw.unembed_w = ptr(dev(sd["patch_unembed.weight"].float().permute(2, 3, 1, 0).reshape(4096, dim), dtype))
        mw.unembed_b = ptr(dev(sd["patch_unembed.bias"], f32))
        mw.dec1_w = ptr(conv64(sd["decoder_conv1.weight"]))
        mw.dec1_b = ptr(dev(sd["decoder_conv1.bias"], f32))
        mw.dec2_w = ptr(conv_to3(sd["decoder_conv2.weight"]))
        mw.dec2_b = ptr(dev(sd["decoder_conv2.bias"], f32))
        if dtype == torch.bfloat16:
            mw.dec2_w16 = ptr(conv_to3_tc(sd["decoder_conv2.weight"]))
            mw.dec2_wst = ptr(conv_to3_stream(sd["decoder_conv2.weight"]))
            mw.dec2_b16 = ptr(bias16(sd["decoder_conv2.bias"]))

        # transformer blocks (q rows and q bias pre-scaled by head_dim^-0.5 = 0.25: exact in fp32 and bf16)
        self.blocks = (_lib.TuBlockWeights * nb)()
        for i in range(nb):
            p = f"{bprefix}{i}."
            bw = self.blocks[i]
            for nm, key in (("ln1_w", "norm1.weight"), ("ln1_b", "norm1.bias"), ("ln2_w", "norm2.weight"), ("ln2_b", "norm2.bias")):
                setattr(bw, nm, ptr(dev(sd[p + key], f32)))
            if resid:
                qw, qb = sd[p + "attn.in_proj_weight"].float().clone(), sd[p + "attn.in_proj_bias"].float().clone()
                pw, pb = sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"]
            else:
                qw, qb = sd[p + "attn.qkv.weight"].float().clone(), sd[p + "attn.qkv.bias"].float().clone()
                pw, pb = sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
                bw.rel_bias = ptr(dev(dense_rel_bias_t(sd[p + "attn.relative_position_bias_table"],
                                                       sd[p + "attn.relative_position_index"]), f32))
            qw[:dim] *= 0.25
            qb[:dim] *= 0.25
            bw.qkv_w, bw.qkv_b = ptr(dev(qw, dtype)), ptr(dev(qb, f32))
            bw.proj_w, bw.proj_b = ptr(dev(pw, dtype)), ptr(dev(pb, f32))
            bw.fc1_w, bw.fc1_b = ptr(dev(sd[p + "mlp.0.weight"], dtype)), ptr(dev(sd[p + "mlp.0.bias"], f32))
            bw.fc2_w, bw.fc2_b = ptr(dev(sd[p + "mlp.2.weight"], dtype)), ptr(dev(sd[p + "mlp.2.bias"], f32))
        mw.blocks = C.cast(self.blocks, C.POINTER(_lib.TuBlockWeights))
        if dtype == torch.bfloat16 and not resid and dim == 128:
            self._pack_fused_stack(sd, mw, nb, dim, bprefix, dev, ptr)
        if dtype == torch.bfloat16 and not resid and dim == 192:
            self._pack_fused_stack192(sd, mw, nb, bprefix, dev, ptr)
        if dtype == torch.bfloat16 and resid and dim == 128:
            self._pack_resid_stack(sd, mw, nb, dim, bprefix, dev, ptr)

        if fast:
            for slot, s in enumerate(SCALES):
                stages = [(0, 2), (2, 2)] if s == 4 else [(0, s)]
                for si, (idx, r) in enumerate(stages):
                    # 64-channel branch: (64 r^2, 64, 3, 3), out channel o = c*r^2 + phase -> [phase][tap][c][ci]
                    w = sd[f"up1.upsamplers.{s}.{idx}.weight"].float()
                    b = sd[f"up1.upsamplers.{s}.{idx}.bias"].float()
                    wp = w.reshape(64, r * r, 64, 3, 3).permute(1, 3, 4, 0, 2).reshape(r * r, 9, 64, 64)
                    mw.up1[slot][si].w = ptr(dev(wp, dtype))
                    mw.up1[slot][si].b = ptr(dev(b.reshape(64, r * r).t().reshape(-1), f32))
                    mw.up1[slot][si].r = r
                    # 3-channel branch: (3 r^2, 3, 3, 3) -> (27, 3 r^2), original out-channel order
                    mw.fin[slot][si].w = ptr(conv_small_in(sd[f"final_upscale.upsamplers.{s}.{idx}.weight"]))
                    mw.fin[slot][si].b = ptr(dev(sd[f"final_upscale.upsamplers.{s}.{idx}.bias"], f32))
                    mw.fin[slot][si].r = r
                if dtype == torch.bfloat16:     # folded last up1 stage + up1_conv (tensor-core path only)
                    idx, r = stages[-1]
                    Wf, bf = fold_up1(sd[f"up1.upsamplers.{s}.{idx}.weight"], sd[f"up1.upsamplers.{s}.{idx}.bias"],
                                      sd["up1_conv.conv.weight"], r)
                    bank, bias = pack_fold_bank(Wf, bf, r)
                    uf = mw.upfold[slot]
                    uf.w, uf.b = ptr(dev(bank, torch.bfloat16)), ptr(dev(bias, f32))
                    uf.ring_w = ptr(dev(Wf.permute(0, 1, 2, 4, 5, 3).reshape(9, 3 * r * r, 25, 64), f32))
                    uf.ring_b = ptr(dev(bf.reshape(9, 3 * r * r), f32))
                    uf.r = r
            mw.up1conv_w = ptr(conv_to3(sd["up1_conv.conv.weight"]))
            if dtype == torch.bfloat16:
                mw.up1conv_w16 = ptr(conv_to3_tc(sd["up1_conv.conv.weight"]))
                mw.up1conv_wst = ptr(conv_to3_stream(sd["up1_conv.conv.weight"]))
                mw.up1conv_b16 = ptr(bias16(None))
            mw.finconv_w = ptr(conv_small_in(sd["final_upscale_conv.weight"]))
            mw.finconv_b = ptr(dev(sd["final_upscale_conv.bias"], f32))
            # host copy of the 3 -> 3 filter: its 84 values travel in the parameters of the fused tail kernel
            wb = torch.cat([sd["final_upscale_conv.weight"].float().permute(2, 3, 1, 0).reshape(-1),
                            sd["final_upscale_conv.bias"].float().reshape(-1)]).cpu().tolist()
            self.host_finconv = (C.c_float * 84)(*wb)
            mw.host_finconv_wb = C.cast(self.host_finconv, C.c_void_p)
        self.struct = mw
        self.dim, self.heads, self.n_blocks = dim, heads, nb

    def _pack_fused_stack192(self, sd, mw, nb, bprefix, dev, ptr):
        """Weights / parameters of all blocks in the order window_stack192_tcgen05.cu consumes them (dim 192, 12 heads).

        Per block 18 slabs [96 n x 64 k]: for each group g of two heads and each K-slab, the rows q | k | v (32 each) of the
        group (q pre-scaled); then proj: 3 slabs [192 n x 64 k]; then the MLP in six chunks of 128 hidden units: fc1 chunk c = 3 slabs
        [128 n x 64 k], fc2 chunk c = 2 slabs [192 n x 64 k], in the order fc1 c0, fc1 c1, then per chunk c: fc2 c, fc1 c+2 (c < 4).
        6912 rows of 64 per block.  Parameters per block: c0 | ln1 w,b | qkv bias in
        the same group-major order | c1 | ln2 w,b | fc1 bias; proj / fc2 biases folded into the offsets c0, c1, c_final.
        """
        f32 = torch.float32
        dim = 192
        slabs, pars, rels = [], [], []
        c = torch.zeros(dim, dtype=torch.float64)
        for i in range(nb):
            p = f"{bprefix}{i}."
            qw, qb = sd[p + "attn.qkv.weight"].float().cpu().clone(), sd[p + "attn.qkv.bias"].float().cpu().clone()
            qw[:dim] *= 0.25
            qb[:dim] *= 0.25
            pw, pb = sd[p + "attn.proj.weight"].float().cpu(), sd[p + "attn.proj.bias"].float().cpu()
            w1, b1 = sd[p + "mlp.0.weight"].float().cpu(), sd[p + "mlp.0.bias"].float().cpu()
            w2, b2 = sd[p + "mlp.2.weight"].float().cpu(), sd[p + "mlp.2.bias"].float().cpu()
            qb_g = []
            for g in range(6):
                rows = torch.cat([torch.arange(s * dim + g * 32, s * dim + g * 32 + 32) for s in range(3)])   # q | k | v of the group
                for ks in range(3):
                    slabs.append(qw[rows, ks * 64:(ks + 1) * 64])
                qb_g.append(qb[rows])
            for ks in range(3):
                slabs.append(pw[:, ks * 64:(ks + 1) * 64])
            # MLP in six chunks of 128 hidden units through two accumulator slots: fc1 c0, fc1 c1, then per chunk c: fc2 c, fc1 c+2
            def fc1_chunk(cc):
                return [w1[cc * 128:(cc + 1) * 128, ks * 64:(ks + 1) * 64] for ks in range(3)]      # [128 n x 64 k]
            def fc2_chunk(cc):
                return [w2[:, cc * 128 + ks * 64: cc * 128 + (ks + 1) * 64] for ks in range(2)]      # [192 n x 64 k]
            slabs += fc1_chunk(0) + fc1_chunk(1)
            for cc in range(6):
                slabs += fc2_chunk(cc)
                if cc + 2 < 6:
                    slabs += fc1_chunk(cc + 2)
            c0 = c.clone()
            c1 = c0 + pb.double()
            c = c1 + b2.double()
            pars += [c0.float(), sd[p + "norm1.weight"].float().cpu(), sd[p + "norm1.bias"].float().cpu(), torch.cat(qb_g),
                     c1.float(), sd[p + "norm2.weight"].float().cpu(), sd[p + "norm2.bias"].float().cpu(), b1]
            rels.append(dense_rel_bias_t(sd[p + "attn.relative_position_bias_table"].cpu(),
                                         sd[p + "attn.relative_position_index"].cpu()))
        pars.append(c.float())
        flat = torch.cat([t.reshape(-1, 64) for t in slabs])
        assert flat.shape[0] == 6912 * nb
        mw.stack_w = ptr(dev(flat, torch.bfloat16))
        mw.stack_p = ptr(dev(torch.cat([t.reshape(-1) for t in pars]), f32))
        mw.stack_rel = ptr(dev(torch.stack([frag_rel_bias(r) for r in rels]), f32))

    def _pack_resid_stack(self, sd, mw, nb, dim, bprefix, dev, ptr):
        """ResidualTransformer, tc/residual_block_tcgen05.cu: per layer the same 24 slabs as the window stack (in_proj 3 n-chunks x 2
        k-slabs with the q rows pre-scaled, out_proj 2, then per hidden half fc1 rows (2 x 2) and fc2 columns (4)) and 1792 parameters:
        c0 = 0 | ln1 w,b | in_proj bias | c1 = out_proj bias | ln2 w,b | fc1 bias | c_final = out_proj bias + fc2 bias.  Every layer
        stores the true residual stream, so the offsets do not run across layers."""
        slabs, pars = [], []
        for i in range(nb):
            p = f"{bprefix}{i}."
            qw, qb = sd[p + "attn.in_proj_weight"].float().cpu().clone(), sd[p + "attn.in_proj_bias"].float().cpu().clone()
            qw[:dim] *= 0.25
            qb[:dim] *= 0.25
            pw, pb = sd[p + "attn.out_proj.weight"].float().cpu(), sd[p + "attn.out_proj.bias"].float().cpu()
            w1, b1 = sd[p + "mlp.0.weight"].float().cpu(), sd[p + "mlp.0.bias"].float().cpu()
            w2, b2 = sd[p + "mlp.2.weight"].float().cpu(), sd[p + "mlp.2.bias"].float().cpu()
            slabs += self._slabs(qw, 3, 2) + self._slabs(pw, 1, 2)
            for h in range(2):
                slabs += self._slabs(w1[h * 256:(h + 1) * 256], 2, 2) + self._slabs(w2[:, h * 256:(h + 1) * 256], 1, 4)
            pars += [torch.zeros(dim), sd[p + "norm1.weight"].float().cpu(), sd[p + "norm1.bias"].float().cpu(), qb,
                     pb, sd[p + "norm2.weight"].float().cpu(), sd[p + "norm2.bias"].float().cpu(), b1,
                     (pb.double() + b2.double()).float()]
        assert len(slabs) == 24 * nb
        mw.stack_w = ptr(dev(torch.stack(slabs).reshape(-1, 64), torch.bfloat16))
        mw.stack_p = ptr(dev(torch.cat([t.reshape(-1) for t in pars]), torch.float32))

    @staticmethod
    def _slabs(w: torch.Tensor, n_chunks: int, k_slabs: int):
        """(n_chunks*128, k_slabs*64) weight -> list of [128 n][64 k] slabs, n-chunk major, k-slab minor."""
        return [w[nc * 128:(nc + 1) * 128, ks * 64:(ks + 1) * 64] for nc in range(n_chunks) for ks in range(k_slabs)]

    def _pack_fused_stack(self, sd, mw, nb, dim, bprefix, dev, ptr):
        """Weights / parameters of all blocks in the order window_stack_tcgen05.cu consumes them.

        Per block 24 slabs: qkv (3 n-chunks x 2 k-slabs, q rows pre-scaled), proj (2 k-slabs), then per hidden half
        h in {0,1}: fc1 rows [256h, 256h+256) (2 n-chunks x 2 k-slabs) and fc2 columns [256h, 256h+256) (4 k-slabs).
        The proj / fc2 biases are not added inside the kernel: they are folded into offset vectors c0 (before LN1),
        c1 (before LN2) and c_final that are added when the residual stream is read.
        """
        f32 = torch.float32
        slabs, pars, rels = [], [], []
        c = torch.zeros(dim, dtype=torch.float64)
        for i in range(nb):
            p = f"{bprefix}{i}."
            qw, qb = sd[p + "attn.qkv.weight"].float().cpu().clone(), sd[p + "attn.qkv.bias"].float().cpu().clone()
            qw[:dim] *= 0.25
            qb[:dim] *= 0.25
            pw, pb = sd[p + "attn.proj.weight"].float().cpu(), sd[p + "attn.proj.bias"].float().cpu()
            w1, b1 = sd[p + "mlp.0.weight"].float().cpu(), sd[p + "mlp.0.bias"].float().cpu()
            w2, b2 = sd[p + "mlp.2.weight"].float().cpu(), sd[p + "mlp.2.bias"].float().cpu()
            slabs += self._slabs(qw, 3, 2) + self._slabs(pw, 1, 2)
            for h in range(2):
                slabs += self._slabs(w1[h * 256:(h + 1) * 256], 2, 2) + self._slabs(w2[:, h * 256:(h + 1) * 256], 1, 4)
            c0 = c.clone()
            c1 = c0 + pb.double()
            c = c1 + b2.double()
            pars += [c0.float(), sd[p + "norm1.weight"].float().cpu(), sd[p + "norm1.bias"].float().cpu(), qb,
                     c1.float(), sd[p + "norm2.weight"].float().cpu(), sd[p + "norm2.bias"].float().cpu(), b1]
            rels.append(dense_rel_bias_t(sd[p + "attn.relative_position_bias_table"].cpu(),
                                         sd[p + "attn.relative_position_index"].cpu()))
        pars.append(c.float())
        assert len(slabs) == 24 * nb
        mw.stack_w = ptr(dev(torch.stack(slabs).reshape(-1, 64), torch.bfloat16))
        mw.stack_p = ptr(dev(torch.cat([t.reshape(-1) for t in pars]), f32))
        mw.stack_rel = ptr(dev(torch.stack([frag_rel_bias(r) for r in rels]), f32))


class CPackedWeights:
    """The same packed weights produced by the C ABI's tu_pack_weights (csrc/pack_weights.cu) instead of the torch code above:
    what a non-Python host does.  Same interface as PackedWeights (`.struct`, `.model`, `.dim`, `.heads`, `.n_blocks`)."""

    def __init__(self, model: str, sd: Dict[str, torch.Tensor], dtype: torch.dtype, device: torch.device):
        lib = _lib.load()
        self.model, self.dtype, self.device = model, dtype, device
        host = {k: (v.detach().to("cpu", torch.int64 if v.dtype == torch.int64 else torch.float32).contiguous()) for k, v in sd.items()}
        arr = (_lib.TuNamedTensor * len(host))()
        self._names = [k.encode() for k in host]
        for i, (k, v) in enumerate(host.items()):
            arr[i].name, arr[i].data, arr[i].numel = self._names[i], v.data_ptr(), v.numel()
        dim = sd["patch_embed.weight"].shape[0]
        pre = "transformer_blocks." if model == "ResidualTransformer" else "window_blocks."
        nb = 1 + max(int(k[len(pre):].split(".")[0]) for k in sd if k.startswith(pre))
        cdt = _lib.TU_BF16 if dtype == torch.bfloat16 else _lib.TU_F32
        nbytes = lib.tu_packed_weights_bytes(_lib.MODEL_IDS[model], dim, nb, cdt)
        if nbytes == 0:
            raise ValueError("tu_packed_weights_bytes: unsupported configuration")
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.packed = _lib.TuPackedModel()
        with torch.cuda.device(device):
            _lib.check(lib.tu_pack_weights(_lib.MODEL_IDS[model], arr, len(host), cdt, self.buf.data_ptr(), nbytes, C.byref(self.packed),
                                           torch.cuda.current_stream(device).cuda_stream))
        self.struct = self.packed.w
        self.dim, self.heads, self.n_blocks = self.struct.dim, self.struct.heads, self.struct.n_blocks
