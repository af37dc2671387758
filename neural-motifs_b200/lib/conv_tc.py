"""3x3 / stride 1 / pad 1 convolution WITH gradients on the tcgen05 path — the missing piece for the detector-training
path (`models/train_detector.py:78-155` trains the VGG backbone and the RPN head, so gradients
must flow through `features` and `rpn_head.conv`, lib/object_detector.py:110-127, 521-531).

STATUS: the gradient formulas below are pinned on the CPU against torch autograd through a torch backend with the same
three primitives (tests/test_conv_tc_walk.py); the kernel backend reuses kernels that are parity-green on the GPU
(implicit-GEMM conv, transposed im2col, bf16x3 GEMM) and is held to fp64 on a B200 by tests/test_detector_train_gpu.py.

Everything is NHWC fp32 at the Function boundary. Three primitives, supplied by a backend:
    conv3x3(x [B,H,W,Ci], wmat [Co, 9*Ci] (k = (kh*3+kw)*Ci + ci), bias|None, relu) -> [B,H,W,Co]
    im2col3_t(x [1,H,W,C]) -> [9*C, H*W]                      (row k = (kh*3+kw)*C + c, zero padding)
    matmul_nt(a [M,K], b [N,K]) -> a @ b^T
and the gradients of y = relu(conv(x, W) + b) follow from them:
    g   = dy * (y > 0)
    dx  = conv3x3(g, wmat_dx)            with wmat_dx[ci, (kh*3+kw)*Co + co] = W[co, ci, 2-kh, 2-kw]
    dW  = sum over images of  matmul_nt(g_b^T [Co, P], im2col3_t(x_b) [9*Ci, P])  -> [Co, 9*Ci] -> [Co, Ci, 3, 3]
    db  = sum over pixels of g
"""
import torch
from torch.autograd import Function


def weight_matrix(w):
    """[Co,Ci,3,3] -> [Co, 9*Ci] with k = (kh*3+kw)*Ci + ci (the K order of the implicit-GEMM kernel)."""
    return w.permute(0, 2, 3, 1).reshape(w.size(0), -1)


def weight_matrix_dx(w):
    """Weight of the data-gradient convolution: [Ci, 9*Co], taps flipped, in/out channels swapped."""
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.size(1), -1)


class KernelBackend(object):
    """The three primitives on csrc/gemm_tc.cu + csrc/maskconv.cu (CUDA tensors only)."""

    def conv3x3(self, x, wmat, bias, relu):
        from lib import tc_ops
        import motifs_cabi as _c
        B, H, W, Ci = x.shape
        Co = wmat.size(0)
        assert Ci % 64 == 0, "the implicit-GEMM conv reads 64-channel K blocks"
        xs = tc_ops.split_rows(x.reshape(-1, Ci))
        ws = tc_ops.split_rows(wmat.contiguous())
        y = torch.empty(B, H, W, Co, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _c.load().mb200_conv3x3_bf16x3(_c.ptr(xs.hi), _c.ptr(xs.lo), _c.ptr(ws.hi), _c.ptr(ws.lo), B, H, W, Ci, Co,
                                                _c.ptr(bias), 1 if relu else 0, _c.ptr(y), None, None, _c.cur_stream())
        _c.check(rc, "mb200_conv3x3_bf16x3")
        return y

    def im2col3_t(self, x):
        from lib import mask_conv
        return mask_conv._im2col3(x.contiguous(), True)            # SplitMat [9C, Pp]

    def matmul_nt(self, a, b):
        """a: fp32 [M,K] given TRANSPOSED as [K,M] rows (so the split kernel transposes it), b: SplitMat [N,Kp]."""
        from lib import tc_ops
        return tc_ops.gemm(tc_ops.split_transposed(a), b)


class TorchBackend(object):
    """The same primitives in plain torch (any device / dtype): the CPU pin of the formulas, never the product path."""

    def conv3x3(self, x, wmat, bias, relu):
        Co, Ci = wmat.size(0), x.size(-1)
        w = wmat.reshape(Co, 3, 3, Ci).permute(0, 3, 1, 2)
        y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, bias, 1, 1).permute(0, 2, 3, 1)
        return torch.relu(y) if relu else y

    def im2col3_t(self, x):
        _, H, W, C = x.shape
        cols = torch.nn.functional.unfold(x.permute(0, 3, 1, 2), 3, padding=1)[0]        # [C*9, P], row = c*9 + tap
        return cols.reshape(C, 9, H * W).permute(1, 0, 2).reshape(9 * C, H * W)           # row = tap*C + c

    def matmul_nt(self, a, b):
        return a.t() @ b.t()                                    # a arrives as [K,M] (see KernelBackend.matmul_nt)


class _Conv3x3(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu, backend):
        y = backend.conv3x3(x, weight_matrix(weight.detach()), bias.detach() if bias is not None else None, relu)
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.relu, ctx.backend, ctx.has_bias = relu, backend, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        be = ctx.backend
        g = dy * (y > 0).to(dy.dtype) if ctx.relu else dy
        g = g.contiguous()
        B, H, W, Ci = x.shape
        Co = weight.size(0)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = be.conv3x3(g, weight_matrix_dx(weight.detach()), None, False)
        if ctx.needs_input_grad[1]:
            acc = None
            for b in range(B):          # per image: bounds the transposed-im2col buffer (9*Ci x H*W bf16 pairs)
                part = be.matmul_nt(g[b].reshape(H * W, Co), be.im2col3_t(x[b:b + 1]))       # [Co, 9*Ci]
                acc = part if acc is None else acc.add_(part)
            dw = acc.reshape(Co, 3, 3, Ci).permute(0, 3, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = g.sum((0, 1, 2))
        return dx, dw, db, None, None


def conv3x3(x, weight, bias=None, relu=True, backend=None):
    """y = relu?(conv2d(x, weight, bias, stride 1, pad 1)) on NHWC fp32 [B,H,W,Ci] -> [B,H,W,Co], with autograd."""
    return _Conv3x3.apply(x, weight, bias, relu, backend if backend is not None else KernelBackend())


def vgg_features_train(x_nchw, convs, cfg, backend=None, stem=None):
    """VGG16 `features` (minus the last pool) with gradients: conv layers through `conv3x3`, 2x2 max-pools through
    torch on the NHWC tensor. `stem(x_nchw, conv) -> NHWC fp32` computes conv1_1 + ReLU (3 input channels do not
    fit the 64-channel K blocks); by default the exact-fp32 stem kernel's forward with an unfold-based weight
    gradient (`_StemConv`). Returns NHWC fp32 [B,H/16,W/16,512]."""
    y = (stem or stem_conv)(x_nchw, convs[0], backend)
    ci = 1
    for v in cfg[1:]:
        if v == 'M':
            y = torch.nn.functional.max_pool2d(y.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()
            continue
        y = conv3x3(y, convs[ci].weight, convs[ci].bias, True, backend)
        ci += 1
    return y


class _StemConv(Function):
    """conv1_1 (3 -> 64) + ReLU. Forward: csrc/stem.cu (exact fp32). Backward: only the weight / bias gradients
    exist (the input is the image): dW = g^T [64, P] x unfold(x) [27, P]^T per image on the bf16x3 GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        import motifs_cabi as _c
        _c.require_cuda(x, weight)
        B, _, H, W = x.shape
        C = weight.size(0)
        dev = x.device
        x = x.contiguous().float()
        yh = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev); yl = torch.empty_like(yh)
        with torch.cuda.device(dev):
            _c.check(_c.load().mb200_conv3x3_stem_split(_c.ptr(x), _c.ptr(weight.detach().contiguous()),
                                                        _c.ptr(bias.detach()), B, H, W, C, 1, _c.ptr(yh), _c.ptr(yl),
                                                        _c.cur_stream()), "mb200_conv3x3_stem_split")
        y = yh.float() + yl.float()
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        from lib import tc_ops
        x, y = ctx.saved_tensors
        g = (dy * (y > 0).float()).contiguous()
        B, H, W, C = g.shape
        acc = None
        for b in range(B):
            cols = torch.nn.functional.unfold(x[b:b + 1], 3, padding=1)[0]               # [27, P], row = c*9 + tap
            part = tc_ops.gemm(tc_ops.split_transposed(g[b].reshape(H * W, C)), tc_ops.split_rows(cols))   # [64, 27]
            acc = part if acc is None else acc.add_(part)
        dw = acc.reshape(C, x.size(1), 3, 3)
        return None, dw, g.sum((0, 1, 2))


def stem_conv(x_nchw, conv, backend=None):
    if backend is not None and not isinstance(backend, KernelBackend):         # CPU pin: plain torch
        return torch.relu(torch.nn.functional.conv2d(x_nchw, conv.weight, conv.bias, 1, 1)).permute(0, 2, 3, 1).contiguous()
    return _StemConv.apply(x_nchw, conv.weight, conv.bias)
