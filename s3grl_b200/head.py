"""Fused scoring head of SIGNNet for the fixed-row flows, evaluation mode.

reference models.py:370-376: `x = operator_diff(torch.cat(xs, -1))` with operator_diff = MLP([(K+1)F', hidden],
batch_norm=True, act_first=True, act='elu', plain_last=False), i.e. Linear -> ELU -> BatchNorm1d (-> dropout,
identity in eval), then models.py:339-346 `_centre_pool_helper` without CCN rows: h[center] * h[center + 1].
`sign_head` does both in one tcgen05 (TF32 in, fp32 accumulate) kernel on the loader's joint matrix
(`JointLoader` batch `.joint`, or a whole epoch); the small link_pred_mlp on [links, hidden] stays in torch.
TF32 rounds the inputs to 10 mantissa bits: results agree with an fp32 reference to ~1e-3 relative — this is the
model head, not the precompute path, whose 1e-5 tolerance is untouched.  No CPU path.
"""
import ctypes as C

import torch

from . import _lib as L

HIDDEN = 256


def fold_batchnorm(bn_weight, bn_bias, running_mean, running_var, eps=1e-5):
    """BatchNorm1d in eval mode as y = x * scale + shift."""
    scale = bn_weight / torch.sqrt(running_var + eps)
    return scale.contiguous(), (bn_bias - running_mean * scale).contiguous()


def sign_head(joint, lin_weight, lin_bias, bn_scale, bn_shift, out=None, stream=None, pool=True):
    """pooled[i] = bn(elu(joint[2i] W^T + b)) * bn(elu(joint[2i+1] W^T + b))  ->  [rows / 2, 256] float32.

    joint [rows, (K+1)F'] float32 CUDA tensor (rows even when pooling), lin_weight [256, (K+1)F'] (torch Linear
    layout), lin_bias / bn_scale / bn_shift [256].  TMA needs 16-byte row strides: the loader's joint matrices
    already have them (padded row stride); any other operand is copied once into a padded buffer.
    pool=False returns h = bn(elu(joint W^T + b)) itself, [rows, 256] (PoS Plus: CCN pooling stays with the caller)."""
    lib = L.lib()
    dev = joint.device
    if dev.type != 'cuda':
        raise RuntimeError("sign_head needs CUDA tensors: there is no CPU path")
    rows, kd = int(joint.shape[0]), int(joint.shape[1])
    if lin_weight.shape != (HIDDEN, kd):
        raise NotImplementedError(f"sign_head serves hidden_channels = {HIDDEN} (got weight {tuple(lin_weight.shape)})")
    if pool and rows % 2:
        raise ValueError("rows must be even for center pooling")

    def tma_ready(t):
        """[r, kd] float32 with unit column stride, 16-byte row stride and base; padded copy otherwise."""
        if (t.dtype == torch.float32 and t.device == dev and t.stride(1) == 1 and t.stride(0) % 4 == 0
                and t.stride(0) >= kd and t.data_ptr() % 16 == 0):
            return t
        buf = torch.zeros((t.shape[0], (kd + 3) // 4 * 4), dtype=torch.float32, device=dev)
        buf[:, :kd].copy_(t)
        return buf[:, :kd]
    joint, lin_weight = tma_ready(joint), tma_ready(lin_weight)
    vecs = [lin_bias, bn_scale, bn_shift]
    if not all(t.dtype == torch.float32 and t.is_contiguous() and t.device == dev and t.numel() == HIDDEN for t in vecs):
        raise ValueError("bias / bn_scale / bn_shift must be contiguous float32 [256] tensors on the joint matrix's device")
    if out is None:
        out = torch.empty((rows // 2 if pool else rows, HIDDEN), dtype=torch.float32, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    p = lambda t: C.c_void_p(t.data_ptr())      # noqa: E731
    with torch.cuda.device(dev):
        L.check(lib.s3_sign_head(p(joint), rows, kd, int(joint.stride(0)), p(lin_weight), int(lin_weight.stride(0)), HIDDEN,
                                 p(lin_bias), p(bn_scale), p(bn_shift), p(out), 1 if pool else 0,
                                 C.c_void_p(st.cuda_stream)), 's3_sign_head')
    return out


_POOL = {'sum': L.POOL_SUM, 'mean': L.POOL_MEAN}


def segment_pool(rows, row_ptr, mode='mean', layout='center', out=None, stream=None):
    """CCN pooling on the GPU (csrc/pool.cu; reference models.py:339-367 with k_heuristic set).

    rows [R, C] float32 CUDA tensor grouped per link by row_ptr [L+1] (row 0 = src, row 1 = dst, rows 2.. = CCN rows).
    layout 'center' -> [L, 2C] = [src * dst | pool(rows 2..)]  (the input of SIGNNet.link_pred_mlp);
    layout 'rows'   -> [L, 3C] = [src | dst | pool(rows 2..)]  (pooled output mode of the precompute path).
    mode 'mean' | 'sum' (k_pool_strategy); anything else raises NotImplementedError as models.py:333 does
    ('concat' needs exactly k_heuristic extra rows per link, which PoS Plus does not guarantee)."""
    if mode not in _POOL:
        raise NotImplementedError(f"Check pool strat: {mode}")
    if layout not in ('center', 'rows'):
        raise ValueError("layout must be 'center' or 'rows'")
    lib = L.lib()
    dev = rows.device
    if dev.type != 'cuda' or rows.dtype != torch.float32 or rows.dim() != 2 or rows.stride(1) != 1:
        raise ValueError("rows must be a [R, C] float32 CUDA tensor with unit column stride (no CPU path)")
    row_ptr = row_ptr.to(device=dev, dtype=torch.int64).contiguous()
    nl, cols = int(row_ptr.numel()) - 1, int(rows.shape[1])
    width = (2 if layout == 'center' else 3) * cols
    if out is None:
        out = torch.empty((nl, width), dtype=torch.float32, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        L.check(lib.s3_segment_pool(C.c_void_p(rows.data_ptr()), int(rows.stride(0)) if rows.shape[0] > 1 else cols, cols,
                                    C.c_void_p(row_ptr.data_ptr()), nl, _POOL[mode],
                                    L.POOL_OUT_CENTER if layout == 'center' else L.POOL_OUT_ROWS,
                                    C.c_void_p(out.data_ptr()), int(out.stride(0)) if nl > 1 else width,
                                    C.c_void_p(st.cuda_stream)), 's3_segment_pool')
    return out


def sign_head_ccn(joint, row_ptr, lin_weight, lin_bias, bn_scale, bn_shift, k_pool_strategy='mean', stream=None):
    """SIGNNet.forward up to link_pred_mlp for the PoS Plus flows (models.py:370-376 + :347-362), evaluation mode:
    h = bn(elu(joint W^T + b)) on the tensor cores (s3_sign_head, pool = 0), then per link
    [h_src * h_dst | mean or sum of its CCN rows' h] on the GPU (s3_segment_pool) -> [L, 512]."""
    h = sign_head(joint, lin_weight, lin_bias, bn_scale, bn_shift, stream=stream, pool=False)
    return segment_pool(h, row_ptr, k_pool_strategy, 'center', stream=stream)
