"""Synthetic networks in CSI-NN2 API terms (lists of tests/shl.py Layer), used by the parity tests,
bench.py and smoke().  Shapes follow the reference's own example graph
(example/c906_mobilenetv1_f16.c:30-1880: 28 convs -- 1 standard 3x3 s2, 13 depthwise 3x3,
13 pointwise 1x1, one 1x1 "fc" 1024->1000 -- 27 standalone relu nodes, global avgpool, softmax);
weights and quantisation parameters are synthetic and seed-fixed: int8
weights uniform in [-127,127] with per-channel symmetric scales, activation qinfo calibrated on
one synthetic image with the float oracle so that every layer uses the int8 range.
"""
from __future__ import annotations

import os

import numpy as np

from shl import (ACT_NONE, DT_F16, DT_INT8, H_ADD, H_CONV, H_FC, H_FLATTEN, H_GAP, H_MAXPOOL, H_RELU,
                 H_SOFTMAX, Layer, Oracle, conv_out_hw)

# (out_channels, stride) of the 13 depthwise-separable blocks
MOBILENET_V1_BLOCKS = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1),
                       (512, 1), (512, 1), (512, 1), (1024, 2), (1024, 1)]


def mobilenet_v1_convs(res=224, width=1.0, classes=1000):
    """[(kind, c_in, c_out, k, stride, pad, depthwise, relu)] in graph order."""
    c = max(8, int(32 * width))
    convs = [("conv", 3, c, 3, 2, 1, False, True)]
    for o, s in MOBILENET_V1_BLOCKS:
        o = max(8, int(o * width))
        convs.append(("conv", c, c, 3, s, 1, True, True))
        convs.append(("conv", c, o, 1, 1, 0, False, True))
        c = o
    convs.append(("gap",))
    convs.append(("conv", c, classes, 1, 1, 0, False, False))
    convs.append(("softmax",))
    return convs


def mobilenet_v1_macs(res=224, width=1.0, classes=1000):
    h = res
    macs = 0
    for spec in mobilenet_v1_convs(res, width, classes):
        if spec[0] == "conv":
            _, ci, co, k, s, p, dw, _ = spec
            h = (h + 2 * p - k) // s + 1
            macs += h * h * co * (1 if dw else ci) * k * k
        elif spec[0] == "gap":
            h = 1
    return macs


def _quant_sym(sigma):
    return float(np.float32(max(3.5 * sigma, 1e-6) / 127.0)), 0


def _quant_pos(sigma):
    return float(np.float32(max(3.5 * sigma, 1e-6) / 255.0)), -128


class NetBuilder:
    """Accumulates layers while propagating one calibration image in float."""

    def __init__(self, dtype, batch, in_chw, seed=0):
        self.dtype, self.batch = dtype, batch
        self.rng = np.random.default_rng(seed)
        self.orc = Oracle()
        self.layers = []
        c, h, w = in_chw
        self.in_shape = (batch, c, h, w)
        self.s_in, self.zp_in = 0.02, 0
        # calibration image = what the int8 input dequantises to
        self.x_calib_q = self.rng.integers(-128, 128, size=(1, c, h, w), dtype=np.int8)
        self.act = [self.x_calib_q.astype(np.float32) * np.float32(self.s_in)]  # float activations by tensor id
        self.q = [(self.s_in, self.zp_in)]
        self.shapes = [self.in_shape]

    # -- helpers ----------------------------------------------------------------------------
    def _push(self, layer, act, q):
        self.layers.append(layer)
        self.act.append(act)
        self.q.append(q)
        self.shapes.append(tuple(layer.out_shape))
        return len(self.layers)  # tensor id of the output

    def conv(self, src, c_out, k, stride, pad, *, depthwise=False, bias=True):
        n, c_in = self.batch, self.shapes[src][1]
        h, w = self.shapes[src][2], self.shapes[src][3]
        oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
        cg = 1 if depthwise else c_in
        kk = cg * k * k
        xf = self.act[src]
        if self.dtype == DT_INT8:
            wq = self.rng.integers(-127, 128, size=(c_out, cg, k, k), dtype=np.int8)
            s_w = (np.float32(np.sqrt(2.0 / kk) / 73.0 / 1.5) * (1.0 + np.arange(c_out) / c_out)).astype(np.float32)
            s_b = (np.float32(self.q[src][0]) * s_w).astype(np.float32)
            bq = self.rng.integers(-2000, 2001, size=(c_out,), dtype=np.int32) if bias else None
            wf = wq.astype(np.float32) * s_w[:, None, None, None]
            bf = None if bq is None else bq.astype(np.float32) * s_b
        else:
            wf16 = (self.rng.standard_normal((c_out, cg, k, k)) * np.sqrt(2.0 / kk)).astype(np.float16)
            bf16 = (0.1 * self.rng.standard_normal(c_out)).astype(np.float16) if bias else None
            wf, bf = wf16.astype(np.float32), None if bf16 is None else bf16.astype(np.float32)
        yf = self.orc.conv2d_f32(xf, wf, bf, (1, c_out, oh, ow), depthwise=depthwise, stride=(stride, stride),
                                 pad=(pad,) * 4)
        if self.dtype == DT_INT8:
            s_out, zp_out = _quant_sym(float(yf.std()))
            layer = Layer(H_CONV, (n, c_out, oh, ow), in0=src, s_out=s_out, zp_out=zp_out, w=wq, b=bq, s_w=s_w,
                          s_b=s_b, stride=(stride, stride), pad=(pad,) * 4, group=c_in if depthwise else 1)
        else:
            s_out, zp_out = 1.0, 0
            layer = Layer(H_CONV, (n, c_out, oh, ow), in0=src, w=wf16, b=bf16, stride=(stride, stride),
                          pad=(pad,) * 4, group=c_in if depthwise else 1)
        return self._push(layer, yf, (s_out, zp_out))

    def relu(self, src):
        yf = np.maximum(self.act[src], 0)
        s_out, zp_out = _quant_pos(float(self.act[src].std())) if self.dtype == DT_INT8 else (1.0, 0)
        if self.dtype == DT_INT8 and os.environ.get("NETS_RELU_KEEPS_QINFO"):
            # experiment knob: the relu node reuses its producer's qinfo (then the fused epilogue is a clamp)
            s_out, zp_out = self.q[src]
        shape = (self.batch,) + self.shapes[src][1:]
        return self._push(Layer(H_RELU, shape, in0=src, s_out=s_out, zp_out=zp_out), yf, (s_out, zp_out))

    def add(self, a, b):
        yf = self.act[a] + self.act[b]
        s_out, zp_out = _quant_sym(float(yf.std())) if self.dtype == DT_INT8 else (1.0, 0)
        shape = (self.batch,) + self.shapes[a][1:]
        return self._push(Layer(H_ADD, shape, in0=a, in1=b, s_out=s_out, zp_out=zp_out), yf, (s_out, zp_out))

    def maxpool(self, src, k, stride, pad):
        x = self.act[src]
        _, c, h, w = x.shape
        oh, ow = conv_out_hw(h, w, k, k, (stride, stride), (pad,) * 4)
        xp = np.full((1, c, h + 2 * pad, w + 2 * pad), -np.inf, np.float32)
        xp[:, :, pad:pad + h, pad:pad + w] = x
        yf = np.full((1, c, oh, ow), -np.inf, np.float32)
        for ky in range(k):
            for kx in range(k):
                yf = np.maximum(yf, xp[:, :, ky:ky + stride * oh:stride, kx:kx + stride * ow:stride])
        q = self.q[src]
        return self._push(Layer(H_MAXPOOL, (self.batch, c, oh, ow), in0=src, s_out=q[0], zp_out=q[1], kernel=(k, k),
                                stride=(stride, stride), pad=(pad,) * 4), yf, q)

    def gap(self, src):
        yf = self.act[src].mean(axis=(2, 3), keepdims=True)
        c = self.shapes[src][1]
        if self.dtype == DT_INT8:
            lo = float(yf.min())
            s_out, zp_out = (_quant_pos(float(yf.max()) / 3.5) if lo >= 0 else _quant_sym(float(np.abs(yf).max()) / 3.5))
        else:
            s_out, zp_out = 1.0, 0
        return self._push(Layer(H_GAP, (self.batch, c, 1, 1), in0=src, s_out=s_out, zp_out=zp_out), yf, (s_out, zp_out))

    def flatten(self, src):
        c = self.shapes[src][1]
        q = self.q[src]
        return self._push(Layer(H_FLATTEN, (self.batch, c), in0=src, s_out=q[0], zp_out=q[1]),
                          self.act[src].reshape(1, c), q)

    def fc(self, src, units):
        c = self.shapes[src][1]
        xf = self.act[src].reshape(1, c)
        if self.dtype == DT_INT8:
            wq = self.rng.integers(-127, 128, size=(units, c), dtype=np.int8)
            s_w = (np.float32(np.sqrt(1.0 / c) / 73.0 / 1.5) * (1.0 + np.arange(units) / units)).astype(np.float32)
            s_b = (np.float32(self.q[src][0]) * s_w).astype(np.float32)
            bq = self.rng.integers(-2000, 2001, size=(units,), dtype=np.int32)
            wf, bf = wq.astype(np.float32) * s_w[:, None], bq.astype(np.float32) * s_b
            yf = xf @ wf.T + bf
            s_out, zp_out = _quant_sym(float(yf.std()))
            layer = Layer(H_FC, (self.batch, units), in0=src, s_out=s_out, zp_out=zp_out, w=wq, b=bq, s_w=s_w, s_b=s_b)
        else:
            wf16 = (self.rng.standard_normal((units, c)) * np.sqrt(1.0 / c)).astype(np.float16)
            bf16 = (0.1 * self.rng.standard_normal(units)).astype(np.float16)
            yf = xf @ wf16.astype(np.float32).T + bf16.astype(np.float32)
            s_out, zp_out = 1.0, 0
            layer = Layer(H_FC, (self.batch, units), in0=src, w=wf16, b=bf16)
        return self._push(layer, yf, (s_out, zp_out))

    def softmax(self, src):
        x = self.act[src].reshape(1, -1)
        e = np.exp(x - x.max())
        yf = (e / e.sum()).reshape(self.act[src].shape)
        s_out, zp_out = (1.0 / 256.0, -128) if self.dtype == DT_INT8 else (1.0, 0)
        shape = (self.batch,) + self.shapes[src][1:]
        return self._push(Layer(H_SOFTMAX, shape, in0=src, s_out=s_out, zp_out=zp_out, axis=1), yf, (s_out, zp_out))

    def input_batch(self, seed=1):
        rng = np.random.default_rng(seed)
        if self.dtype == DT_INT8:
            return rng.integers(-128, 128, size=self.in_shape, dtype=np.int8)
        return rng.standard_normal(self.in_shape).astype(np.float16)


def mobilenet_v1(dtype=DT_INT8, batch=1, res=224, width=1.0, classes=1000, seed=0, softmax=True) -> NetBuilder:
    nb = NetBuilder(dtype, batch, (3, res, res), seed)
    t = 0
    for spec in mobilenet_v1_convs(res, width, classes):
        if spec[0] == "conv":
            _, ci, co, k, s, p, dw, relu = spec
            t = nb.conv(t, co, k, s, p, depthwise=dw)
            if relu:
                t = nb.relu(t)
        elif spec[0] == "gap":
            t = nb.gap(t)
        elif softmax:
            t = nb.softmax(t)
    return nb


def resnet50(dtype=DT_INT8, batch=1, res=224, width=1.0, classes=1000, seed=0) -> NetBuilder:
    """torchvision ResNet-50 v1.5 shapes (stride on the 3x3 of each downsampling bottleneck);
    not part of the reference tree (BASELINE.json configs[4])."""
    nb = NetBuilder(dtype, batch, (3, res, res), seed)
    nb.zp_in = -7 if dtype == DT_INT8 else 0  # "asymmetric" activations
    nb.q[0] = (nb.s_in, nb.zp_in)
    nb.act[0] = (nb.x_calib_q.astype(np.float32) - np.float32(nb.zp_in)) * np.float32(nb.s_in)
    w0 = max(8, int(64 * width))
    t = nb.relu(nb.conv(0, w0, 7, 2, 3))
    t = nb.maxpool(t, 3, 2, 1)
    c_in = w0
    for stage, (blocks, mid) in enumerate([(3, 64), (4, 128), (6, 256), (3, 512)]):
        mid = max(8, int(mid * width))
        for b in range(blocks):
            stride = 2 if (b == 0 and stage > 0) else 1
            y = nb.relu(nb.conv(t, mid, 1, 1, 0))
            y = nb.relu(nb.conv(y, mid, 3, stride, 1))
            y = nb.conv(y, mid * 4, 1, 1, 0)
            sc = t
            if b == 0:
                sc = nb.conv(t, mid * 4, 1, stride, 0)
            t = nb.relu(nb.add(y, sc))
            c_in = mid * 4
    t = nb.gap(t)
    t = nb.flatten(t)
    t = nb.fc(t, classes)
    return nb


def oracle_forward(nb: NetBuilder, x: np.ndarray, all_values: bool = False):
    """Whole-network result by chaining the oracle's per-op restatements (int8: the exact
    contract, so the product must match it bit for bit; fp16: f32 math rounded to f16 per layer)."""
    from shl import H_AVGPOOL, H_CONV_RELU, H_CONV_RELU6, H_DWCONV, H_RELU6, H_RESHAPE, ACT_RELU, ACT_RELU6
    orc = nb.orc
    vals = [x]
    qs = [(nb.s_in, nb.zp_in)]
    i8 = nb.dtype == DT_INT8
    for l in nb.layers:
        a = vals[l.in0]
        s_in, zp_in = qs[l.in0]
        if l.kind in (H_CONV, H_CONV_RELU, H_CONV_RELU6, H_DWCONV):
            dw = l.group > 1 and l.w.shape[1] == 1 and l.group == a.shape[1]
            act = {H_CONV_RELU: ACT_RELU, H_CONV_RELU6: ACT_RELU6}.get(l.kind, ACT_NONE)
            if i8:
                y = orc.conv2d_i8(a, l.w, l.b, tuple(l.out_shape), depthwise=dw, stride=l.stride, pad=l.pad,
                                  dilation=l.dilation, group=1 if dw else l.group, s_in=s_in, zp_in=zp_in,
                                  s_w=l.s_w, s_b=l.s_b, s_out=l.s_out, zp_out=l.zp_out, act=act, zp_w=l.zp_w)
            else:
                y = orc.conv2d_f32(a.astype(np.float32), l.w.astype(np.float32),
                                   None if l.b is None else l.b.astype(np.float32), tuple(l.out_shape),
                                   depthwise=dw, stride=l.stride, pad=l.pad, dilation=l.dilation,
                                   group=1 if dw else l.group, act=act).astype(np.float16)
        elif l.kind == H_FC:
            if i8:
                y = orc.fc_i8(a.reshape(a.shape[0], -1), l.w, l.b, s_in=s_in, zp_in=zp_in, s_w=l.s_w, s_b=l.s_b,
                              s_out=l.s_out, zp_out=l.zp_out, zp_w=l.zp_w)
            else:
                y = (a.reshape(a.shape[0], -1).astype(np.float32) @ l.w.astype(np.float32).T +
                     l.b.astype(np.float32)).astype(np.float16)
        elif l.kind in (H_RELU, H_RELU6):
            act = ACT_RELU if l.kind == H_RELU else ACT_RELU6
            if i8:
                y = orc.relu_i8(a, act, s_in, zp_in, l.s_out, l.zp_out)
            else:
                y = np.maximum(a, 0) if act == ACT_RELU else np.clip(a, 0, 6)
        elif l.kind == H_ADD:
            b = vals[l.in1]
            if i8:
                y = orc.add_i8(a, b, s_in, zp_in, qs[l.in1][0], qs[l.in1][1], l.s_out, l.zp_out)
            else:
                y = (a.astype(np.float32) + b.astype(np.float32)).astype(np.float16)
        elif l.kind in (H_MAXPOOL, H_AVGPOOL, H_GAP):
            kernel = (a.shape[2], a.shape[3]) if l.kind == H_GAP else tuple(l.kernel)
            stride = (1, 1) if l.kind == H_GAP else tuple(l.stride)
            pad = (0, 0, 0, 0) if l.kind == H_GAP else tuple(l.pad)
            if i8:
                y = orc.pool_i8(a, tuple(l.out_shape), avg=l.kind != H_MAXPOOL, kernel=kernel, stride=stride,
                                pad=pad, count_include_pad=l.count_include_pad, s_in=s_in, zp_in=zp_in,
                                s_out=l.s_out, zp_out=l.zp_out)
            else:
                assert l.kind == H_GAP, "fp16 oracle chain covers global avgpool only"
                acc = np.zeros(a.shape[:2], np.float32)
                for yy in range(a.shape[2]):
                    for xx in range(a.shape[3]):
                        acc = acc + a[:, :, yy, xx].astype(np.float32)
                y = (acc / np.float32(a.shape[2] * a.shape[3])).astype(np.float16).reshape(l.out_shape)
        elif l.kind == H_SOFTMAX:
            a2 = a.reshape(a.shape[0], -1)
            if i8:
                y = orc.softmax_i8(a2, s_in, zp_in, l.s_out, l.zp_out).reshape(l.out_shape)
            else:
                z = a2.astype(np.float32)
                e = np.exp((z - z.max(axis=1, keepdims=True)).astype(np.float64))
                y = (e / e.sum(axis=1, keepdims=True)).astype(np.float16).reshape(l.out_shape)
        elif l.kind in (H_FLATTEN, H_RESHAPE):
            y = a.reshape(l.out_shape)
        else:
            raise NotImplementedError(l.kind)
        vals.append(np.ascontiguousarray(y))
        qs.append((l.s_out, l.zp_out))
    return vals if all_values else vals[-1]  # vals[i] = tensor i (0 = the input, i = output of layer i - 1)
