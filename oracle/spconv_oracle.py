"""TEST INFRASTRUCTURE ONLY - never imported by the product package (tests/, __graft_entry__.smoke() and bench.py's CPU legs only).

CPU restatement (torch, fp32 / fp64) of the sparse 3-D convolution semantics the reference's sparse U-Nets rely on (SURVEY 8f
next-1).  The reference delegates to spconv; the algorithm restated here is the one vendored in the tree
(mmdet3d/ops/spconv, v1):

  index pairs      include/spconv/geometry.h:25-86 (getValidOutPos: out = (in + pad - delta) / stride, kernel offset
                   index = (dz*kY + dy)*kX + dx), :200-252 (getIndicePairsConv), :254-297 (getIndicePairsSubM)
  gather-GEMM      include/spconv/spconv_ops.h:95-260 (indiceConv: out[o] += in[i] @ W[k] for every pair (i, o) of offset k;
                   `inverse` swaps the roles of the pair columns)
  layer glue       mmdet3d/ops/spconv/conv.py:110-206, modules.py:113-127
  U-Net            mmdet3d/models/middle_encoders/sparse_unet.py:114-207, 369-413, 470-505; ops/sparse_block.py:126-143

Restated through DENSE tensors, which is what makes the restatement independent of any rulebook code: a sparse convolution equals
torch's dense conv3d (cross-correlation, zero padding) of the scattered input, read at the active output cells; the active
output cells of a strided SparseConv are the cells whose receptive field holds an active input; a SubM conv reads only the input
cells; SparseInverseConv equals conv_transpose3d with the couple conv's stride / padding read at the couple conv's INPUT cells.

Pinning: tests/test_oracle_vs_reference.py checks every function below against the reference's own vendored spconv v1 (C++ sources
compiled unmodified by oracle/build_ref.py into oracle/_ref, Python layers imported unmodified through oracle/ref_shim.py) on random
sparse tensors incl. whole SimpleSparseUNet / VirtualVoxelMixer forwards, and tests/golden/spconv_*.npz hold outputs of that
reference for the GPU box ("parity pinned" on the in-tree v1; spconv 2.2.3, which docs/overall_instructions.md:28 installs instead,
is not in the tree - its arithmetic contract is the same one).  Output row ORDER of a strided SparseConv is implementation-defined in spconv (hash insertion order); here and in the
CUDA path it is lexicographic (b,z,y,x), and comparisons with the reference are made per coordinate."""
import numpy as np
import torch
import torch.nn.functional as F


def _t3(v):
    return [int(x) for x in v] if isinstance(v, (list, tuple)) else [int(v)] * 3


def conv_output_size(in_shape, ksize, stride, padding):
    """ops.py:20-31 (dilation 1)"""
    return [(i + 2 * p - (k - 1) - 1) // s + 1 for i, k, s, p in zip(in_shape, ksize, stride, padding)]


def _dense(feats, coors, batch_size, shape):
    d = torch.zeros((batch_size, feats.shape[1]) + tuple(shape), dtype=feats.dtype)
    c = coors.long()
    d[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = feats
    return d


def _gather(dense, coors):
    c = coors.long()
    return dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]


def out_coors(coors, batch_size, in_shape, ksize, stride, padding):
    """Active output cells of SparseConv3d, lexicographic (b,z,y,x) (geometry.h:200-252: every valid out position of every input)."""
    ks, st, pd = _t3(ksize), _t3(stride), _t3(padding)
    occ = _dense(torch.ones((coors.shape[0], 1)), coors, batch_size, in_shape)
    hit = F.conv3d(occ, torch.ones((1, 1) + tuple(ks)), stride=st, padding=pd) > 0.5
    out = torch.nonzero(hit[:, 0])   # row-major = lexicographic
    return out.int()


def neighbour_table(in_coors, out_coors_, batch_size, in_shape, ksize, stride, padding):
    """nbr[o][k] = input row at out*stride - pad + delta_k or -1 (the pairs of geometry.h:60-67 regrouped by output row)."""
    ks, st, pd = _t3(ksize), _t3(stride), _t3(padding)
    grid = torch.full((batch_size,) + tuple(in_shape), -1, dtype=torch.long)
    ci = in_coors.long()
    grid[ci[:, 0], ci[:, 1], ci[:, 2], ci[:, 3]] = torch.arange(ci.shape[0])
    co = out_coors_.long()
    kv = ks[0] * ks[1] * ks[2]
    nbr = torch.full((co.shape[0], kv), -1, dtype=torch.long)
    k = 0
    for dz in range(ks[0]):
        for dy in range(ks[1]):
            for dx in range(ks[2]):
                z = co[:, 1] * st[0] - pd[0] + dz
                y = co[:, 2] * st[1] - pd[1] + dy
                x = co[:, 3] * st[2] - pd[2] + dx
                ok = (z >= 0) & (z < in_shape[0]) & (y >= 0) & (y < in_shape[1]) & (x >= 0) & (x < in_shape[2])
                v = grid[co[:, 0], z.clamp(0, in_shape[0] - 1), y.clamp(0, in_shape[1] - 1), x.clamp(0, in_shape[2] - 1)]
                nbr[:, k] = torch.where(ok, v, torch.full_like(v, -1))
                k += 1
    return nbr.int()


def _w_conv(weight):
    """(kD,kH,kW,in,out) [conv.py:97-98] -> torch conv3d weight (out,in,kD,kH,kW)"""
    return weight.permute(4, 3, 0, 1, 2).contiguous()


def sparse_conv(feats, coors, batch_size, in_shape, weight, stride, padding, bias=None):
    """SparseConv3d.forward (conv.py:110-206, subm=False).  Returns (out_feats, out_coors, out_shape)."""
    ks = list(weight.shape[:3])
    st, pd = _t3(stride), _t3(padding)
    oc = out_coors(coors, batch_size, in_shape, ks, st, pd)
    dense = F.conv3d(_dense(feats, coors, batch_size, in_shape), _w_conv(weight), bias=bias, stride=st, padding=pd)
    return _gather(dense, oc), oc, conv_output_size(in_shape, ks, st, pd)


def subm_conv(feats, coors, batch_size, in_shape, weight, bias=None):
    """SubMConv3d.forward: outputs only at the input cells, out[o] = sum_k W[k] . in[o - k//2 + delta_k].  The layer's `padding` /
    `stride` arguments do not reach the index generation: getIndicePair forces stride 1 and padding k//2 for SubM
    (spconv_ops.h:74-78; spconv 2.x's generate_subm_conv_inds takes no padding either), so e.g. VirtualVoxelMixer.conv_out
    (kernel 3, padding 0; sparse_unet.py:457-467) is a centred 3x3x3 SubM conv."""
    ks = list(weight.shape[:3])
    dense = F.conv3d(_dense(feats, coors, batch_size, in_shape), _w_conv(weight), bias=bias, stride=1, padding=[k // 2 for k in ks])
    return _gather(dense, coors)


def inverse_conv(feats, coors, batch_size, out_shape_of_couple, couple_in_coors, couple_in_shape, weight, stride, padding, bias=None):
    """SparseInverseConv3d.forward (conv.py:147-153, 189-192): the couple conv's pairs with in / out swapped, i.e.
    out[i] = sum over (o, k) with o*stride - pad + delta_k == i of feats[o] @ W[k], at the couple conv's input cells."""
    ks = list(weight.shape[:3])
    st, pd = _t3(stride), _t3(padding)
    dense = _dense(feats, coors, batch_size, out_shape_of_couple)
    opad = [couple_in_shape[d] - ((out_shape_of_couple[d] - 1) * st[d] - 2 * pd[d] + ks[d]) for d in range(3)]
    extra = [max(0, -o) for o in opad]
    wt = weight.permute(3, 4, 0, 1, 2).contiguous()   # conv_transpose3d weight: (in, out, kD, kH, kW)
    full = F.conv_transpose3d(dense, wt, bias=None, stride=st, padding=pd, output_padding=[max(0, o) for o in opad])
    if any(extra):
        full = full[:, :, :couple_in_shape[0], :couple_in_shape[1], :couple_in_shape[2]]
    out = _gather(full, couple_in_coors)
    if bias is not None:
        out = out + bias
    return out


def indice_conv(feats, nbr, weight, scale=None, shift=None, residual=None, relu=False):
    """What sstb200_spconv_forward computes, from a neighbour table (fp64 accumulation): the per-launch checker."""
    kv = nbr.shape[1]
    w = weight.reshape(kv, weight.shape[-2], weight.shape[-1]).double()
    out = torch.zeros((nbr.shape[0], w.shape[2]), dtype=torch.float64)
    f = feats.double()
    for k in range(kv):
        idx = nbr[:, k].long()
        m = idx >= 0
        if m.any():
            out[m] += f[idx[m]] @ w[k]
    if scale is not None:
        out = out * scale.double()
    if shift is not None:
        out = out + shift.double()
    if residual is not None:
        out = out + residual.double()
    if relu:
        out = out.clamp_min(0)
    return out.float()


# ----------------------------------------------------------------------------------------------------------------------
# the U-Nets, driven by a reference-layout state dict
# ----------------------------------------------------------------------------------------------------------------------
class _T:
    def __init__(self, feats, coors, shape, batch_size, pairs=None):
        self.f, self.c, self.shape, self.b = feats, coors, list(shape), batch_size
        self.pairs = {} if pairs is None else pairs   # indice_key -> dict(in_coors, in_shape, out_coors, out_shape, stride, padding, subm)

    def with_f(self, f):
        return _T(f, self.c, self.shape, self.b, self.pairs)


def _bn_eval(x, sd, prefix, eps):
    return (x - sd[prefix + "running_mean"]) / torch.sqrt(sd[prefix + "running_var"] + eps) * sd[prefix + "weight"] + sd[prefix + "bias"]


def _conv(x, sd, prefix, kind, key, stride=1, padding=0):
    """one spconv layer with the reference's indice_key reuse (conv.py:147-172)."""
    w = sd[prefix + "weight"]
    if kind == "subm":
        x.pairs.setdefault(key, dict(subm=True))
        return _T(subm_conv(x.f, x.c, x.b, x.shape, w), x.c, x.shape, x.b, x.pairs)
    if kind == "conv":
        assert key not in x.pairs
        f, oc, oshape = sparse_conv(x.f, x.c, x.b, x.shape, w, stride, padding)
        x.pairs[key] = dict(in_coors=x.c, in_shape=x.shape, out_shape=oshape, stride=_t3(stride), padding=_t3(padding), subm=False)
        return _T(f, oc, oshape, x.b, x.pairs)
    rec = x.pairs[key]   # inverse
    f = inverse_conv(x.f, x.c, x.b, rec["out_shape"], rec["in_coors"], rec["in_shape"], w, rec["stride"], rec["padding"])
    return _T(f, rec["in_coors"], rec["in_shape"], x.b, x.pairs)


def _block(x, sd, prefix, kind, key, eps, stride=1, padding=0, act=True):
    """make_sparse_convmodule(order=conv,norm,act) in eval mode: `prefix`0 = conv, `prefix`1 = BN"""
    y = _conv(x, sd, prefix + "0.", kind, key, stride, padding)
    f = _bn_eval(y.f, sd, prefix + "1.", eps)
    return y.with_f(F.relu(f) if act else f)


def _basic_block(x, sd, prefix, key, eps):
    """SparseBasicBlock.forward (sparse_block.py:126-143), eval mode"""
    out = _conv(x, sd, prefix + "conv1.", "subm", key)
    out = out.with_f(F.relu(_bn_eval(out.f, sd, prefix + "bn1.", eps)))
    out = _conv(out, sd, prefix + "conv2.", "subm", key)
    out = out.with_f(_bn_eval(out.f, sd, prefix + "bn2.", eps))
    return out.with_f(F.relu(out.f + x.f))


def sparse_unet_forward(sd, voxel_feats, coors, batch_size, sparse_shape, encoder_channels, encoder_paddings, decoder_channels,
                        decoder_paddings, eps=1e-3, mixer_out=False, return_multiscale=False):
    """SimpleSparseUNet.forward / VirtualVoxelMixer.forward (sparse_unet.py:369-413 / 470-505) in eval mode over a reference-layout
    state dict.  Returns (voxel_feats, voxel_coors[, decoder feature list])."""
    sd = {k: v.float() for k, v in sd.items() if torch.is_tensor(v)}
    x = _block(_T(voxel_feats.float(), coors.int(), sparse_shape, batch_size), sd, "conv_input.", "subm", "subm1", eps, padding=1)
    enc = []
    for i, blocks in enumerate(encoder_channels):
        for j in range(len(blocks)):
            pad = tuple(encoder_paddings[i])[j]
            p = f"encoder_layers.encoder_layer{i + 1}.{j}."
            if i != 0 and j == 0:
                x = _block(x, sd, p, "conv", f"spconv{i + 1}", eps, stride=2, padding=pad)
            else:
                x = _block(x, sd, p, "subm", f"subm{i + 1}", eps, padding=pad)
        enc.append(x)
    n = len(decoder_channels)
    x = enc[-1]
    ms = []
    for i in range(n, 0, -1):
        lat = _basic_block(enc[i - 1], sd, f"lateral_layer{i}.", f"subm{i}", eps)
        cat = lat.with_f(torch.cat((x.f, lat.f), dim=1))
        merged = _block(cat, sd, f"merge_layer{i}.", "subm", f"subm{i}", eps)
        cm = merged.f.shape[1]
        red = cat.f.view(cat.f.shape[0], cm, -1).sum(dim=2)
        y = cat.with_f(merged.f + red)
        if i != 1:
            x = _block(y, sd, f"upsample_layer{i}.", "inv", f"spconv{i}", eps)
        else:
            x = _block(y, sd, f"upsample_layer{i}.", "subm", "subm1", eps)
        ms.append(x)
    if mixer_out:
        x = _block(x, sd, "conv_out.", "subm", "out_conv", eps, padding=0)
    if return_multiscale:
        return x.f, x.c, [(m.f, m.c) for m in ms]
    return x.f, x.c


# small U-Net / mixer configurations of the golden fixtures (oracle/make_golden.py spconv_fixture)
SP_UNET = dict(in_channels=8, sparse_shape=[9, 32, 32], order=('conv', 'norm', 'act'), norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01),
               base_channels=8, output_channels=16, encoder_channels=((8,), (8, 8, 8), (16, 16, 16), (16, 16, 16)),
               encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
               decoder_channels=((16, 16, 16), (16, 16, 8), (8, 8, 8), (8, 8, 8)), decoder_paddings=((1, 1), (1, 0), (0, 0), (0, 1)))
SP_MIXER = dict(in_channels=8, sparse_shape=[8, 24, 24], norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01), base_channels=8,
                output_channels=12, encoder_channels=((8,), (8, 8), (8, 8)), encoder_paddings=((1,), (1, 1), (1, 1)),
                decoder_channels=((8, 8, 8), (8, 8, 8), (8, 8, 8)), decoder_paddings=((1, 1), (1, 1), (1, 1)))

# the mixer of the config-5 fixture (tests/golden/fsdv2_front_mixer.npz)
FSDV2_MIXER = dict(in_channels=32, sparse_shape=[12, 160, 160], norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01), base_channels=8,
                   output_channels=16, encoder_channels=((8,), (8, 8), (16, 16)), encoder_paddings=((1,), (1, 1), (1, 1)),
                   decoder_channels=((16, 16, 8), (8, 8, 8), (8, 8, 8)), decoder_paddings=((1, 1), (1, 1), (1, 1)))


def synth_sparse(seed, batch_size, shape, n_per_sample, channels, clustered=True):
    """Deterministic sparse tensor for tests: unique (b,z,y,x) int32 rows in random order + fp32 features."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(batch_size):
        if clustered:   # a few blobs, so that 3x3x3 neighbourhoods are populated like a voxelised surface
            centres = torch.stack([torch.randint(0, s, (6,), generator=g) for s in shape], 1).float()
            pick = centres[torch.randint(0, 6, (n_per_sample * 2,), generator=g)]
            pts = pick + torch.randn((n_per_sample * 2, 3), generator=g) * torch.tensor([1.0, 3.0, 3.0])
            zyx = torch.stack([pts[:, d].round().clamp(0, shape[d] - 1) for d in range(3)], 1).long()
        else:
            zyx = torch.stack([torch.randint(0, s, (n_per_sample * 2,), generator=g) for s in shape], 1)
        zyx = torch.unique(zyx, dim=0)
        zyx = zyx[torch.randperm(zyx.shape[0], generator=g)[:n_per_sample]]
        rows.append(torch.cat([torch.full((zyx.shape[0], 1), b), zyx], 1))
    coors = torch.cat(rows).int()
    coors = coors[torch.randperm(coors.shape[0], generator=g)]
    feats = torch.randn((coors.shape[0], channels), generator=g)
    return feats, coors
