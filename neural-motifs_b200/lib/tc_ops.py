"""Host side of the bf16x3 tensor-core path (csrc/gemm_tc.cu): operand splitting with a
per-parameter cache, `linear_tc` (drop-in for F.linear with autograd) and the VGG feature
extractor built from the implicit-GEMM 3x3 convolution.

The arithmetic replaced is nn.Linear / nn.Conv2d of the reference (cuBLAS / cuDNN fp32):
fc6/fc7 `lib/object_detector.py:129-138`, `lib/rel_model.py:360-374,439-448`, post_lstm /
rel_compress `lib/rel_model.py:377,390,503,524`, VGG16 features `lib/object_detector.py:110-127`.
"""
import weakref

import torch
from torch.autograd import Function

import motifs_cabi as _c


def _round_up(x, m):
    return (x + m - 1) // m * m


class SplitMat(object):
    """(hi, lo) bf16 pair of a logical [rows, K] fp32 matrix, row pitch Kp (multiple of 64)."""
    __slots__ = ("hi", "lo", "rows", "K", "Kp")

    def __init__(self, hi, lo, rows, K, Kp):
        self.hi, self.lo, self.rows, self.K, self.Kp = hi, lo, rows, K, Kp


def split_rows(x):
    """x [R,K] fp32 CUDA (row stride arbitrary, unit column stride) -> SplitMat of x."""
    _c.require_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32
    if x.stride(1) != 1:
        x = x.contiguous()
    R, K = x.shape
    Kp = _round_up(K, 64)
    hi = torch.empty(R, Kp, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(R, Kp, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _c.check(_c.load().mb200_split_bf16(_c.ptr(x), R, K, x.stride(0), Kp, _c.ptr(hi), _c.ptr(lo), _c.cur_stream()),
                 "mb200_split_bf16")
    return SplitMat(hi, lo, R, K, Kp)


def split_transposed(x):
    """x [R,C] fp32 CUDA -> SplitMat of x^T: logical [C, R], pitch round_up(R, 64)."""
    _c.require_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32
    if x.stride(1) != 1:
        x = x.contiguous()
    R, Cc = x.shape
    Rp = _round_up(R, 64)
    hi = torch.empty(Cc, Rp, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(Cc, Rp, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _c.check(_c.load().mb200_split_transpose_bf16(_c.ptr(x), R, Cc, x.stride(0), Rp, _c.ptr(hi), _c.ptr(lo),
                                                      _c.cur_stream()), "mb200_split_transpose_bf16")
    return SplitMat(hi, lo, Cc, R, Rp)


# Optional per-launch profiling (bench.py's roofline leg): when PROFILE is a list, every tensor-core
# launch appends (kind, algorithmic_flops, start_event, end_event) recorded on the launching stream.
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_end(kind, flops, ev0):
    if ev0 is None:
        return
    ev1 = torch.cuda.Event(enable_timing=True)
    ev1.record()
    PROFILE.append((kind, float(flops), ev0, ev1))


def gemm(A, B, bias=None, relu=False, want_f32=True, want_split=False, out=None):
    """C = A @ B^T (+bias)(relu) with A [M,K], B [N,K] SplitMats. Returns fp32 [M,N] and/or a SplitMat
    of C (pitch round_up(N,64), zero padded) per the flags. `out`: contiguous fp32 [M,N] to write into."""
    assert A.Kp == B.Kp, (A.Kp, B.Kp)
    M, N = A.rows, B.rows
    dev = A.hi.device
    if out is not None:
        assert want_f32 and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == M * N
        C = out
    else:
        C = torch.empty(M, N, dtype=torch.float32, device=dev) if want_f32 else None
    Cs = None
    if want_split:
        Np = _round_up(N, 64)
        alloc = torch.zeros if Np != N else torch.empty
        Cs = SplitMat(alloc(M, Np, dtype=torch.bfloat16, device=dev), alloc(M, Np, dtype=torch.bfloat16, device=dev),
                      M, N, Np)
    lib = _c.load()
    ws_n = lib.mb200_gemm_workspace_floats(M, N, A.Kp)
    ws = torch.empty(ws_n, dtype=torch.float32, device=dev) if ws_n > 0 else None
    ev0 = _prof_begin()
    with torch.cuda.device(dev):
        rc = lib.mb200_gemm_bf16x3(_c.ptr(A.hi), _c.ptr(A.lo), _c.ptr(B.hi), _c.ptr(B.lo), M, N, A.Kp,
                                   _c.ptr(bias), 1 if relu else 0, _c.ptr(C), N,
                                   _c.ptr(Cs.hi) if Cs else None, _c.ptr(Cs.lo) if Cs else None,
                                   Cs.Kp if Cs else 0, _c.ptr(ws), _c.cur_stream())
    _c.check(rc, "mb200_gemm_bf16x3")
    _prof_end("gemm", 2.0 * M * N * A.K, ev0)
    if want_f32 and want_split:
        return C, Cs
    return C if want_f32 else Cs


# Weight gradients dW = dY^T X run on the MN-major kernel (csrc/gemm_mn.cu) straight from the row-major (hi, lo) pairs —
# dY split ONCE per backward (shared with dX = dY W), X's split kept from forward — instead of two transposing split
# passes per layer (round 1: split_transpose_kernel 1.04 ms + part of split_kernel's 0.84 ms per step).
# MOTIFS_GEMM_MN=0 restores the K-major route (A/B runs).
GEMM_MN = __import__("os").environ.get("MOTIFS_GEMM_MN", "1") == "1"


def gemm_mn(At, Bt, out=None):
    """C[M,N] = At^T @ Bt with At a SplitMat of [K, M] (rows = the reduction index) and Bt of [K, N]:
    the weight-gradient product dW = dY^T X straight from the row-major activations (csrc/gemm_mn.cu)."""
    assert At.rows == Bt.rows, (At.rows, Bt.rows)
    K, M, N = At.rows, At.K, Bt.K
    dev = At.hi.device
    if out is not None:
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == M * N
        C = out
    else:
        C = torch.empty(M, N, dtype=torch.float32, device=dev)
    lib = _c.load()
    ws_n = lib.mb200_gemm_mn_workspace_floats(M, N, K)
    ws = torch.empty(ws_n, dtype=torch.float32, device=dev) if ws_n > 0 else None
    ev0 = _prof_begin()
    with torch.cuda.device(dev):
        rc = lib.mb200_gemm_bf16x3_mn(_c.ptr(At.hi), _c.ptr(At.lo), At.Kp, _c.ptr(Bt.hi), _c.ptr(Bt.lo), Bt.Kp, M, N, K,
                                      _c.ptr(C), N, _c.ptr(ws), _c.cur_stream())
    _c.check(rc, "mb200_gemm_bf16x3_mn")
    _prof_end("gemm", 2.0 * M * N * K, ev0)
    return C


# ------------------------------------------------------------------ weight split cache
# key (id(param), kind) -> (weakref to the parameter, version tuple, value). The entry dies with the parameter
# (weakref callback), so a new model whose parameters reuse a dead model's id() / caching-allocator address
# can never hit the dead model's splits, and rebuilt models do not leak their fc6 splits (0.8 GB each).
_cache = {}
WEIGHT_EPOCH = 0     # bumped by optimizers that update parameters through raw pointers (lib/fused_optim.py)
LOAD_EPOCH = 0       # bumped by load_state_dict hooks and by invalidate_all(): applies to frozen parameters too


def bump_weight_epoch():
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1


def invalidate_all():
    """Every cached split is stale from now on (trainable AND frozen parameters). `load_state_dict` of RelModel /
    ObjectDetector calls this through a post hook. A raw `param.data.copy_(...)` AFTER a forward changes neither
    `_version` nor `data_ptr` (the checkpoint idiom of models/train_rels.py:86-95 runs before the first forward and
    is safe): callers that write `.data` later must call this (or clear_cache())."""
    global LOAD_EPOCH
    LOAD_EPOCH += 1


def install_load_hook(module):
    """load_state_dict on `module` invalidates the split cache (state-dict loads copy in place)."""
    module.register_load_state_dict_post_hook(lambda mod, incompatible: invalidate_all())


def _lookup(owner, key, ver, maker, src):
    hit = _cache.get(key)
    if hit is not None and hit[0]() is owner and hit[1] == ver:
        return hit[2]
    val = maker(src.detach())

    def _evict(ref, key=key):
        cur = _cache.get(key)
        if cur is not None and cur[0] is ref:
            del _cache[key]
    _cache[key] = (weakref.ref(owner, _evict), ver, val)
    return val


def _version_of(t, shape):
    return (t.data_ptr(), t._version, tuple(shape), WEIGHT_EPOCH if t.requires_grad else 0, LOAD_EPOCH)


def preset_rows_split(param, hi, lo):
    """An optimizer that has just written the parameter AND its bf16 pair (lib/fused_optim.FlatSGD, presplit) hands the pair
    over: valid until the next weight / load epoch. `hi`, `lo`: [N, K] bf16 views with K % 64 == 0."""
    param._mb200_presplit = ((WEIGHT_EPOCH, LOAD_EPOCH, param.data_ptr()), SplitMat(hi, lo, param.size(0), param.size(1), param.size(1)))


def _cached(param, kind, maker):
    """Split copies of a parameter are rebuilt only when the parameter changes (optimizer steps bump `_version`
    or the weight epoch); frozen weights are split once per load."""
    if kind == "rows":
        ps = getattr(param, "_mb200_presplit", None)
        if ps is not None and ps[0] == (WEIGHT_EPOCH, LOAD_EPOCH, param.data_ptr()):
            return ps[1]
    owner = param._base if param._base is not None else param     # views (w.view(out, -1)) are temporaries: key on the base
    key = (id(owner), kind) if owner is param else \
        (id(owner), kind, param.storage_offset(), tuple(param.shape), tuple(param.stride()))
    return _lookup(owner, key, _version_of(owner, param.shape), maker, param)


def _cached_view(base, tag, view, maker):
    """Like _cached, for a view (slice) of the flat parameter `base`; `tag` names the slice."""
    return _lookup(base, (id(base), tag), _version_of(base, view.shape), maker, view)


def weight_split(weight):          # [N,K] -> B operand of  x @ W^T
    return _cached(weight, "rows", split_rows)


def weight_split_t(weight):        # [N,K] -> W^T as [K,N]: B operand of  dY @ W
    return _cached(weight, "cols", split_transposed)


def clear_cache():
    _cache.clear()


# ------------------------------------------------------------------ direct gradient writes
class DirectGradState(object):
    """Attached (as `param._mb200_direct`) by lib/fused_optim.FlatSGD to parameters whose `.grad` is a view of a
    flat buffer the fused step has just zeroed. A weight-gradient GEMM may then write its result straight into
    `.grad` (or into the slice of it that belongs to a view of the parameter) instead of materialising a
    temporary for autograd's AccumulateGrad to add: for fc6 that pass alone re-reads and re-writes 1.2 GB.
    `written`: slices already written this step; `dirty`: AccumulateGrad has run since the last step (then
    `.grad` is no longer known to be zero and the ordinary path is taken). The data-parallel chunk countdown
    stays on the AccumulateGrad hook: torch runs it once per backward for every parameter in the graph, after
    all of its uses have been differentiated, also when every use returned no gradient tensor — announcing a
    parameter from the op that wrote it would be too early for a parameter used twice."""
    __slots__ = ("written", "dirty")

    def __init__(self):
        self.written, self.dirty = set(), False

    def reset(self):
        self.written.clear()
        self.dirty = False


DIRECT_GRADS = True       # switch for A/B runs (bench.py --no-direct-grads)


def direct_grad_target(param, view=None):
    """The tensor a gradient GEMM may overwrite for `param` (or for `view`, a contiguous slice of it), else None."""
    st = getattr(param, "_mb200_direct", None)
    if not DIRECT_GRADS or st is None or st.dirty or param.grad is None or not param.grad.is_contiguous():
        return None
    if view is None:
        if st.written:
            return None
        st.written.add("all")
        return param.grad
    if "all" in st.written or not view.is_contiguous():
        return None
    off = view.storage_offset() - param.storage_offset()
    if off < 0 or off + view.numel() > param.numel():
        return None
    key = (off, view.numel())
    if any(k[0] < off + view.numel() and off < k[0] + k[1] for k in st.written):
        return None                                   # overlaps a slice already written this step
    st.written.add(key)
    return param.grad.view(-1)[off:off + view.numel()].view(view.shape)


class _LinearTC(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        xs = split_rows(x.detach())
        y = gemm(xs, weight_split(weight), bias=bias.detach() if bias is not None else None)
        ctx.has_bias = bias is not None
        if GEMM_MN:         # keep the bf16 pair of x (same bytes as x): it is the B operand of the weight-gradient GEMM
            ctx.save_for_backward(weight, xs.hi, xs.lo)
            ctx.xdims = (xs.rows, xs.K, xs.Kp)
        else:
            ctx.save_for_backward(weight, x)
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        gx = gw = gb = None
        gys = split_rows(gy) if (ctx.needs_input_grad[0] or (GEMM_MN and ctx.needs_input_grad[1])) else None
        if GEMM_MN:
            weight, xhi, xlo = ctx.saved_tensors
        else:
            weight, x = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            gx = gemm(gys, weight_split_t(weight))                              # [M,N] x [K,N]^T -> [M,K]
        if ctx.needs_input_grad[1]:
            tgt = direct_grad_target(weight)
            if GEMM_MN:
                gw = gemm_mn(gys, SplitMat(xhi, xlo, *ctx.xdims), out=tgt)        # dY^T X, no transposed copies
            else:
                gw = gemm(split_transposed(gy), split_transposed(x.detach()), out=tgt) # [N,M] x [K,M]^T -> [N,K]
            if tgt is not None:
                gw = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb


class _MatmulTC(Function):
    """y = x @ w with w [K,N] a view of the flat parameter `base` (highway-LSTM weight layout,
    alternating_highway_lstm.py:212-221). `tag` identifies the view for the split cache."""

    @staticmethod
    def forward(ctx, x, w, base, tag):
        xs = split_rows(x.detach())
        y = gemm(xs, _cached_view(base, (tag, "T"), w, split_transposed))
        if GEMM_MN:
            ctx.save_for_backward(w, xs.hi, xs.lo)
            ctx.xdims = (xs.rows, xs.K, xs.Kp)
        else:
            ctx.save_for_backward(w, x)
        ctx.base, ctx.tag = base, tag
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        gx = gw = None
        gys = split_rows(gy) if (ctx.needs_input_grad[0] or (GEMM_MN and ctx.needs_input_grad[1])) else None
        if GEMM_MN:
            w, xhi, xlo = ctx.saved_tensors
        else:
            w, x = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            gx = gemm(gys, _cached_view(ctx.base, (ctx.tag, "R"), w, split_rows))   # gy @ w^T
        if ctx.needs_input_grad[1]:
            tgt = direct_grad_target(ctx.base, w)
            if GEMM_MN:
                gw = gemm_mn(SplitMat(xhi, xlo, *ctx.xdims), gys, out=tgt)         # x^T @ gy
            else:
                gw = gemm(split_transposed(x.detach()), split_transposed(gy), out=tgt)
            if tgt is not None:
                gw = None          # written into the slice of base.grad that belongs to this view
        return gx, gw, None, None


def matmul_tc(x, w, base, tag):
    return _MatmulTC.apply(x, w, base, tag)


def linear_tc(x, weight, bias=None):
    """F.linear(x, weight, bias) on the tcgen05 path (fp32 in / fp32 out, autograd aware)."""
    _c.require_cuda(x, weight)
    shp = x.shape
    y = _LinearTC.apply(x.reshape(-1, shp[-1]), weight, bias)
    return y.reshape(*shp[:-1], weight.size(0))


def linear_tc_nograd(x_split, weight, bias=None, relu=False, want_f32=True, want_split=False):
    """Inference-side linear on an already split input; ReLU fused; may emit the split output so
    that chained layers (fc6 -> fc7) never materialise fp32 activations."""
    return gemm(x_split, weight_split(weight), bias=bias, relu=relu, want_f32=want_f32, want_split=want_split)


# ------------------------------------------------------------------ VGG16 features (frozen, forward only)
VGG16_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512]


def _conv_weight_split(conv_weight):
    def mk(w):
        O, I = w.size(0), w.size(1)
        dev = w.device
        lib = _c.load()
        w = w.contiguous()
        if I == 3:
            hi = torch.empty(O, 64, dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi)
            with torch.cuda.device(dev):
                _c.check(lib.mb200_stem_weight_split(_c.ptr(w), O, _c.ptr(hi), _c.ptr(lo), _c.cur_stream()), "stem_weight")
            return SplitMat(hi, lo, O, 27, 64)
        Ip = _round_up(I, 64)
        hi = torch.empty(O, 9 * Ip, dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi)
        with torch.cuda.device(dev):
            _c.check(lib.mb200_conv_weight_split(_c.ptr(w), O, I, Ip, _c.ptr(hi), _c.ptr(lo), _c.cur_stream()), "conv_weight")
        return SplitMat(hi, lo, O, 9 * I, 9 * Ip)
    return _cached(conv_weight, "conv", mk)


def conv3x3_relu(xs, B, H, W, Cin, conv, want_f32=False, want_split=True, relu=True):
    """xs: (hi, lo) NHWC bf16 tensors [B,H,W,Cin]; conv: nn.Conv2d(3x3, pad 1, stride 1). Bias (if the
    module has one) and ReLU (`relu`) fused. Returns (y_f32 NHWC or None, (yhi, ylo) or None)."""
    wsp = _conv_weight_split(conv.weight)
    Cout = conv.weight.size(0)
    dev = xs[0].device
    y = torch.empty(B, H, W, Cout, dtype=torch.float32, device=dev) if want_f32 else None
    yh = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev) if want_split else None
    yl = torch.empty_like(yh) if want_split else None
    ev0 = _prof_begin()
    with torch.cuda.device(dev):
        rc = _c.load().mb200_conv3x3_bf16x3(_c.ptr(xs[0]), _c.ptr(xs[1]), _c.ptr(wsp.hi), _c.ptr(wsp.lo), B, H, W, Cin,
                                            Cout, _c.ptr(conv.bias.detach()) if conv.bias is not None else None,
                                            1 if relu else 0,
                                            _c.ptr(y), _c.ptr(yh), _c.ptr(yl), _c.cur_stream())
    _c.check(rc, "mb200_conv3x3_bf16x3")
    _prof_end("conv3x3", 2.0 * B * H * W * Cout * 9 * Cin, ev0)
    return y, ((yh, yl) if want_split else None)


def maxpool2(xs, B, H, W, C):
    dev = xs[0].device
    yh = torch.empty(B, H // 2, W // 2, C, dtype=torch.bfloat16, device=dev)
    yl = torch.empty_like(yh)
    with torch.cuda.device(dev):
        _c.check(_c.load().mb200_maxpool2_nhwc_split(_c.ptr(xs[0]), _c.ptr(xs[1]), B, H, W, C, _c.ptr(yh), _c.ptr(yl),
                                                     _c.cur_stream()), "mb200_maxpool2_nhwc_split")
    return (yh, yl)


def vgg_features_forward(x, convs, want_last_split=False):
    """x [B,3,H,W] fp32 NCHW; convs: the 13 nn.Conv2d of VGG16 `features` minus the last max-pool
    (load_vgg, lib/object_detector.py:623-633). Returns conv5_3+ReLU as NHWC fp32 [B,H/16,W/16,512].
    Forward only (the detector is frozen in train_rels.py:51-52 and its output is detached)."""
    _c.require_cuda(x)
    x = x.contiguous().float()
    B, _, H, W = x.shape
    dev = x.device
    lib = _c.load()
    C = convs[0].weight.size(0)
    if C == 64 and convs[0].weight.size(1) == 3:
        # stem: exact-fp32 direct convolution straight to the NHWC bf16 pair (csrc/stem.cu)
        xh = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev); xl = torch.empty_like(xh)
        w0 = convs[0].weight.detach().contiguous()
        with torch.cuda.device(dev):
            _c.check(lib.mb200_conv3x3_stem_split(_c.ptr(x), _c.ptr(w0), _c.ptr(convs[0].bias.detach()), B, H, W, C, 1,
                                                  _c.ptr(xh), _c.ptr(xl), _c.cur_stream()), "mb200_conv3x3_stem_split")
        xs = (xh, xl)
    else:
        # generic stem: explicit im2col (K = 27 -> 64) + plain GEMM with fused bias/ReLU
        a_hi = torch.empty(B * H * W, 64, dtype=torch.bfloat16, device=dev); a_lo = torch.empty_like(a_hi)
        with torch.cuda.device(dev):
            _c.check(lib.mb200_im2col3_split(_c.ptr(x), B, H, W, _c.ptr(a_hi), _c.ptr(a_lo), _c.cur_stream()), "im2col3")
        w0 = _conv_weight_split(convs[0].weight)
        cur = gemm(SplitMat(a_hi, a_lo, B * H * W, 27, 64), w0, bias=convs[0].bias.detach(), relu=True,
                   want_f32=False, want_split=True)
        xs = (cur.hi.view(B, H, W, C), cur.lo.view(B, H, W, C))
    ci = 1
    out = None
    for li, v in enumerate(VGG16_CFG[1:], start=1):
        if v == 'M':
            xs = maxpool2(xs, B, H, W, C)
            H, W = H // 2, W // 2
            continue
        last = ci == len(convs) - 1
        out, xs = conv3x3_relu(xs, B, H, W, C, convs[ci], want_f32=last, want_split=(not last) or want_last_split)
        C = v
        ci += 1
    return out, (xs if want_last_split else None)
