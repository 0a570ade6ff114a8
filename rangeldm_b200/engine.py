"""Plan compiler: turns a `UNet2DModel` / `AutoencoderKL` module tree into a flat program of
librldm kernel launches (`rldm_op` records, include/rldm.h) over statically allocated HBM buffers.

Data layout in HBM (per plan, allocated once):
  * boundary tensors keep the reference layout (B, C, W, H) fp32 (`ldm/dataset.py:230`);
  * every internal activation is channels-last (B, W, H, C): the fp32 "residual stream" written by
    conv epilogues, and the fp16 tensor-core operand written by `prep` (GroupNorm-apply + SiLU +
    skip-concat + nearest-upsample + cast in one pass);
  * weights are repacked once to [tap][Cout][Cin] fp16 (tap = kW*3 + kH), projections of an
    attention block are stacked into one [3C][C] GEMM, all `time_emb_proj` rows into one matrix;
  * GroupNorm moments for all layers live in one double arena cleared by a single memset per forward.

The program is replayed with one `rldm_run` call (no host work between launches) and is CUDA-graph
capturable.  Reference semantics followed: SURVEY.md App. A.1/A.2, `ldm/utils.py:40-58,107-116`,
`vae/sgm/modules/diffusionmodules/model.py:342-362,1024-1057`.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import RldmOp

# debug switch used by the GPU tests to run the CUDA-core restatement of the conv (never the default)
CONV_KIND = _lib.OP_CONV_TC

# Tensor-core operand precision, chosen PER LAYER (`terms` of rldm_conv_tc_ex):
#   3 "fp16x3": split-fp16 -- activations and weights are carried as hi+lo fp16 pairs and every K step issues
#               Xh*Wh + Xl*Wh + Xh*Wl into the fp32 TMEM accumulator (~22-bit operands).
#   2 "fp16x2": activations single fp16, weights hi+lo: Xh*Wh + Xh*Wl.  No low-order activation plane exists: the
#               producer pass writes, and the conv reads, half the operand bytes; 2/3 of the tensor work.
#   1 "fp16":   plain fp16 operands (1 MMA per K step).
# fp16 operands carry 11 bits.  Measured against the fp32 oracle (profiles/precision_pareto2_r2.json, C3 shapes,
# max|a-b|/max|b|, tolerance 1e-3):
#   * ONE UNet forward: all fp16 5.3e-4, full-resolution level at fp16x3 (rest fp16) 2.3e-4, all fp16x3 3e-6 -- the
#     full-resolution level (K = 1152, last layers before conv_out) sets the error;
#   * a 20-step DPM-Solver++ trajectory damps each step's eps error: all fp16 1.25e-4 on the final latent;
#   * the VAE decoder (~25 convolutions, nothing damps it): fp16x2 1.0e-3 and fp16 1.5e-3 on an N(0,1) latent, fp16x3 2e-5.
# Hence two UNet profiles: a lone `unet(x, t)` call keeps its full-resolution level at fp16x3 (every single forward
# < 5e-4), the fused sampler's trajectory programs run all fp16 (trajectories < 2e-4, 1574 vs 1756 us per forward), and
# the VAE stays at fp16x3.  The latency-bound lower UNet levels gain little from fewer MMAs but nothing from more.
#   RLDM_PRECISION              UNet levels 1..n (both profiles)            default fp16
#   RLDM_PRECISION_TOP          UNet full-resolution level, lone forwards   default fp16x3
#   RLDM_PRECISION_TOP_SAMPLER  UNet full-resolution level, trajectories    default fp16
#   RLDM_PRECISION_VAE          VAE decoder / encoder                       default fp16x3
_TERMS = {"fp16x3": 3, "fp16x2": 2, "fp16": 1}


def _terms_env(name, default):
    v = os.environ.get(name, default)
    if v not in _TERMS:
        raise ValueError(f"{name} must be one of {sorted(_TERMS)}, got {v!r}")
    return _TERMS[v]


PRECISION = _terms_env("RLDM_PRECISION", "fp16")
PRECISION_TOP = _terms_env("RLDM_PRECISION_TOP", "fp16x3")
PRECISION_TOP_SAMPLER = _terms_env("RLDM_PRECISION_TOP_SAMPLER", "fp16")
PRECISION_VAE = _terms_env("RLDM_PRECISION_VAE", "fp16x3")
# GroupNorm moments accumulated in the conv epilogue (RLDM_FUSE_STATS=0 forces the separate rldm_gn_stats pass)
FUSE_STATS = os.environ.get("RLDM_FUSE_STATS", "1") != "0"
# ResnetBlock2D.conv_shortcut folded into conv2's launch (RLDM_FUSE_SHORTCUT=0: separate 1x1 launch + fp32 residual)
FUSE_SHORTCUT = os.environ.get("RLDM_FUSE_SHORTCUT", "1") != "0"
FUSE_CONV_OUT = os.environ.get("RLDM_FUSE_CONV_OUT", "1") != "0"
# RLDM_FUSE_PREP=1 (experiment, default off): the convolutions of the small layers (fewer 128 x 128 tiles than SMs: UNet
# levels 1..n) produce their own fp16 operand from the fp32 stream inside the kernel (rldm_conv_tc_fused) instead of
# reading the output of a rldm_prep launch: 118 graph nodes per UNet forward instead of 168.  Bit-identical
# (tests/test_fused_gpu.py) but measured SLOWER on B200 (C3, batch 8: 174 vs 216 images/s): the one 192-thread CTA per
# SM needs 10-22 k cycles for its slice (every output-channel tile of the same pixels repeats it, and six warps cannot
# hide the load -> SiLU -> store chain), against ~5 us for the whole rldm_prep launch including its kernel boundary
# (scripts/own_operand_probe.py).
FUSE_PREP = os.environ.get("RLDM_FUSE_PREP", "0") == "1"
# The small-layer convolutions EMIT the next GroupNorm's operand (rldm_conv_tc_emit): where a convolution's output is
# consumed through GroupNorm (+ SiLU) by another convolution on the same grid (norm2 -> conv2 of a ResnetBlock2D, the
# GroupNorm of the next block), its epilogue completes the (image, group) moments inside a thread-block cluster that
# covers whole images, normalises its own rows and writes the fp16 operand -- no rldm_prep launch, no kernel boundary.
# librldm only accepts layers whose images fit RLDM_EMIT_MAXCLM tiles of 128 pixels (default 1: level 3 of the C3 UNet,
# 17 of the 66 prep launches of a forward: 230.3 -> 232.7 images/s).  With clusters that also span the M tiles of larger
# images (RLDM_EMIT_MAXCLM=2 / 8: 28 / 40 launches fewer) the 8-CTA clusters are placed later than the 2-4-CTA clusters
# they replace and the whole step loses (218.7 / 209 images/s, scripts/emit_probe.py).  RLDM_EMIT_PREP=0: never.
EMIT_PREP = os.environ.get("RLDM_EMIT_PREP", "1") != "0"
# Upsample2D of the layers the role-swapped kernel takes (the VAE decoder's): the nearest-2x upsampling is folded into
# the convolution as four 2x2 phase convolutions over the low-resolution operand (RLDM_FOLD_UPSAMPLE=0: upsample in the
# prep pass, then the 3x3 convolution on four times the pixels).
FOLD_UPSAMPLE = os.environ.get("RLDM_FOLD_UPSAMPLE", "1") != "0"
# RLDM_FUSE_LEVELS=1 (experiment, default off): runs of small consecutive ops (UNet levels 1..n) compiled into ONE
# persistent launch each (csrc/fused_levels.cu).  Correct (tests/test_fused_gpu.py) but measured SLOWER on B200: a
# grid-wide barrier costs 2.0-2.3 us against ~3 us for a PDL kernel boundary, and every convolution needs two of them
# (K-slice reduction, then GroupNorm moments): UNet forward 1.87 -> 2.97 ms (profiles/fused_levels_timeline_r2.txt).
FUSE_LEVELS = os.environ.get("RLDM_FUSE_LEVELS", "0") == "1"


def _require_cuda_device(dev, what):
    """Plans hold device pointers; compiling one for a CPU model is only allowed as a DRY RUN
    (RLDM_DRYRUN=1: host-logic tests count ops and check shapes; `run()` then refuses)."""
    if dev.type != "cuda" and os.environ.get("RLDM_DRYRUN") != "1":
        raise RuntimeError(f"{what} must be on a CUDA device: rangeldm_b200 has no CPU fallback")


class Program:
    """A flat list of `rldm_op` records over statically allocated buffers.  `ops` is what the builders emit (one record
    per kernel-level operation); `finalize()` compiles maximal runs of small ops into fused persistent launches
    (`rldm_fused_*`, include/rldm.h) and `exec_ops` / `arr` is what `run()` replays."""

    def __init__(self, device, fuse=True):
        self.device = device
        self.ops = []
        self.launches = []      # kernel launches of each op (parallel to `ops`)
        self.keep = []          # every tensor the program points into
        self._free = {}         # (nbytes) -> [tensor]
        self.arr = None
        self.exec_ops, self.exec_launches = None, None
        self.fuse = fuse        # False for programs that run CONCURRENTLY with others (a fused launch needs all SMs)
        self._fused_handles = []
        self.no_reuse = os.environ.get("RLDM_NOFREE") == "1"   # debugging: keep every intermediate alive
        self.taps = []          # (module, Act) pairs recorded by the Builder (debugging / tests)

    @property
    def n_launch(self):
        return sum(self.exec_launches if self.exec_launches is not None else self.launches)

    # ---- buffers ---------------------------------------------------------------------------
    def alloc(self, shape, dtype=torch.float32, before_op=None):
        """before_op=k: the buffer will be WRITTEN by op k, which has already been emitted (an epilogue that produces
        the operand of a later op): only buffers that were free before op k was appended may be reused."""
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        lst = self._free.get(nbytes)
        raw = None
        if lst:
            if before_op is None:
                raw = lst.pop()[0]
            else:
                for k in range(len(lst) - 1, -1, -1):
                    if lst[k][1] <= before_op:
                        raw = lst.pop(k)[0]
                        break
        if raw is None:
            raw = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.keep.append(raw)
        return raw.view(dtype).view(*shape)

    def free(self, t):
        if t is None or self.no_reuse:
            return
        raw = t.view(-1).view(torch.uint8)
        self._free.setdefault(raw.numel(), []).append((raw, len(self.ops)))     # free from op index len(ops) on

    def hold(self, t):
        self.keep.append(t)
        return t

    # ---- ops -------------------------------------------------------------------------------
    def add(self, kind, i=(), f=(), p=(), n=0, launches=1):
        op = RldmOp()
        op.kind = kind
        for k, v in enumerate(i):
            op.i[k] = int(v)
        for k, v in enumerate(f):
            op.f[k] = float(v)
        for k, v in enumerate(p):
            op.p[k] = None if v is None else (v if isinstance(v, int) else v.data_ptr())
        op.n = int(n)
        self.ops.append(op)
        self.launches.append(launches)
        return op

    def append(self, op, launches=1):
        """Append an already built record (a patched copy of another program's op)."""
        self.ops.append(op)
        self.launches.append(launches)

    def extend(self, other):
        """Append another program's ops (sharing its buffers)."""
        self.ops.extend(other.ops)
        self.launches.extend(other.launches)
        self.keep.append(other)

    def _segments(self):
        """Maximal runs [i, j) of consecutive ops the fused-levels kernel accepts (at least two ops, one of them a
        convolution: a lone op gains nothing from a persistent launch)."""
        if not (FUSE_LEVELS and self.fuse and CONV_KIND == _lib.OP_CONV_TC):
            return []
        lib = _lib.lib()
        segs, i, n = [], 0, len(self.ops)
        while i < n:
            j = i
            while j < n and lib.rldm_fused_supported(ctypes.byref(self.ops[j])):
                j += 1
            if j - i >= 2 and any(self.ops[k].kind == _lib.OP_CONV_TC for k in range(i, j)):
                segs.append((i, j))
            i = max(j, i + 1)
        return segs

    def finalize(self):
        self._release_fused()
        segs = self._segments()
        lib = _lib.lib() if segs else None
        ws, ws_bytes = None, 0
        if segs and self.device.type == "cuda":
            for i, j in segs:
                arr = (RldmOp * (j - i))(*self.ops[i:j])
                ws_bytes = max(ws_bytes, int(lib.rldm_fused_ws_bytes(arr, j - i)))
            ws = self.hold(torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=self.device))
        ex, el, src, pos = [], [], [], 0
        for i, j in segs:
            ex.extend(self.ops[pos:i]); el.extend(self.launches[pos:i]); src.extend((k, k + 1) for k in range(pos, i))
            op = RldmOp()
            op.kind = _lib.OP_FUSED
            op.n = j - i
            if self.device.type == "cuda":
                arr = (RldmOp * (j - i))(*self.ops[i:j])
                h = ctypes.c_void_p()
                with torch.cuda.device(self.device):
                    _lib.check(lib.rldm_fused_create(arr, j - i, ws.data_ptr(), ws_bytes, ctypes.byref(h)))
                self._fused_handles.append(h)
                op.p[0] = h.value
            ex.append(op); el.append(1); src.append((i, j))
            pos = j
        ex.extend(self.ops[pos:]); el.extend(self.launches[pos:]); src.extend((k, k + 1) for k in range(pos, len(self.ops)))
        self.exec_ops, self.exec_launches = ex, el
        self.exec_src = src     # exec op k stands for ops[src[k][0]:src[k][1]]
        self.arr = (RldmOp * len(ex))(*ex)
        return self

    def _release_fused(self):
        hs, self._fused_handles = self._fused_handles, []
        for h in hs:
            try:
                _lib.lib().rldm_fused_destroy(h)
            except Exception:      # interpreter shutdown
                pass

    def __del__(self):
        self._release_fused()

    def run(self):
        if self.device.type != "cuda":
            raise RuntimeError("dry-run program (built on CPU) cannot execute: no CPU fallback")
        if self.arr is None:
            self.finalize()
        _lib.check(_lib.lib().rldm_run(self.arr, len(self.exec_ops), _lib.stream_ptr()))


class Operand:
    """fp16 tensor-core operand pair (hi, lo) in the W-padded layout (B, W+2, H, C).  While `src` is set the operand
    has not been produced yet: the consuming convolution produces it itself (small-layer kernel) or `Builder` emits
    the rldm_prep launch when it turns out that it cannot."""
    __slots__ = ("hi", "lo", "src")

    def __init__(self, hi, lo, src=None):
        self.hi, self.lo, self.src = hi, lo, src

    def __getitem__(self, k):
        return (self.hi, self.lo)[k]


class Act:
    """Channels-last fp32 activation (B, W, H, C).  `stats` = arena slice [B][C/2][2] when the producing conv already
    accumulated the channel-pair moments of this tensor in its epilogue."""
    __slots__ = ("t", "B", "W", "H", "C", "stats", "producer")

    def __init__(self, t, B, W, H, C, stats=None, producer=None):
        self.t, self.B, self.W, self.H, self.C, self.stats = t, B, W, H, C, stats
        self.producer = producer      # the rldm_conv_tc record that writes this tensor (it may emit a consumer's operand)


def _is_identity_attn(m):
    return m is None or not hasattr(m, "to_q")


class Builder:
    """Emits ops for the building blocks shared by the UNet and the VAE."""

    def __init__(self, prog, batch, max_gn=4096, groups=32, cache=None, terms_of=None):
        self.cache = cache if cache is not None else {}      # packed weights shared between plans of a model
        # terms_of(W) -> 1 | 2 | 3: operand precision of a convolution whose operand grid is W columns wide
        self.terms_of = terms_of if terms_of is not None else (lambda W: PRECISION)
        self.pg = prog
        self.B = batch
        self.groups = groups
        self.gn_arena = prog.hold(torch.zeros(max_gn * batch * groups * 2, dtype=torch.float64, device=prog.device))   # 16 MB at batch 8
        self.gn_used = 0
        self.memset_op = prog.add(_lib.OP_MEMSET, p=(self.gn_arena,), n=0)
        self.temb = None         # (tensor (B,T), T)
        self.temb_rows = {}      # id(resnet) -> row offset

    def finish(self):
        self.memset_op.n = self.gn_used * 8

    # ---- weights ---------------------------------------------------------------------------
    def terms(self, W):
        t = self.terms_of(W)
        if CONV_KIND == _lib.OP_CONV_REF and t == 2:         # the CUDA-core restatement knows 1 and 3 only
            t = 3
        return t

    def _cached(self, key, make):
        key = key + (str(self.pg.device),)
        v = self.cache.get(key)
        if v is None:
            v = self.cache[key] = make()
        self.pg.keep.append(v)
        return v

    def f32(self, t):
        return self._cached(("f32", id(t)), lambda: t.detach().to(self.pg.device, torch.float32).contiguous())

    def _planes(self, w, terms):
        """fp32 [taps][Cout][Cin] -> fp16 [planes][taps][Cout][Cin]; plane 1 = residual of the fp16 rounding
        (present for terms >= 2)."""
        hi = w.to(torch.float16)
        if terms < 2:
            return self.pg.hold(hi.contiguous())
        lo = (w - hi.float()).to(torch.float16)
        return self.pg.hold(torch.cat([hi, lo], 0).contiguous())

    def pack_conv(self, conv, terms):
        def make():
            w = conv.weight.detach().to(self.pg.device, torch.float32)       # (Cout, Cin, kW, kH)
            co, ci, k0, k1 = w.shape
            return self._planes(w.permute(2, 3, 0, 1).reshape(k0 * k1, co, ci), terms)
        wt = self._cached(("conv", id(conv.weight), min(terms, 2)), make)
        b = self.f32(conv.bias) if conv.bias is not None else None
        return wt, b

    def pack_linear(self, lins, terms):
        def make():
            w = torch.cat([l.weight.detach().to(self.pg.device, torch.float32) for l in lins], 0)
            b = torch.cat([l.bias.detach().to(self.pg.device, torch.float32) for l in lins], 0)
            return self._planes(w[None], terms), b.contiguous()
        return self._cached(("lin", min(terms, 2)) + tuple(id(l.weight) for l in lins), make)

    def alloc_half(self, shape, terms):
        """(hi, lo) fp16 operand pair; lo only exists for a split-fp16 x3 consumer."""
        return Operand(self.pg.alloc(shape, torch.float16), self.pg.alloc(shape, torch.float16) if terms == 3 else None)

    def free_half(self, pair):
        self.pg.free(pair[0])
        self.pg.free(pair[1])

    # ---- primitive emitters ----------------------------------------------------------------
    def stats_slot(self, groups):
        """`groups` (sum, sum^2) slots per image from the per-forward arena (zeroed by one fill per forward)."""
        n = self.B * groups * 2
        off = self.gn_used
        self.gn_used += n
        assert self.gn_used <= self.gn_arena.numel(), "GroupNorm arena exhausted"
        return self.gn_arena[off:off + n]

    def gn_stats(self, x0, x1, groups):
        sums = self.stats_slot(groups)
        c1 = x1.C if x1 is not None else 0
        self.pg.add(_lib.OP_GN_STATS, i=(x0.C, c1, self.B, x0.W * x0.H, groups),
                    p=(x0.t, x1.t if x1 is not None else None, sums))
        return sums

    def prep(self, x0, x1=None, norm=None, silu=False, up=1, circular=True, also_raw=False, terms=3, raw_terms=3,
             defer=False):
        """-> fp16 operand pair in the W-padded layout (B, W*up + 2, H*up, C0+C1); also_raw=True returns a
        second pair holding the un-normalised input (1x1 shortcut operand) written by the same launch.  `terms` /
        `raw_terms`: precision of the consuming convolutions (the lo plane is written for 3 only).
        defer=True: no launch is emitted yet -- the operands carry their recipe (`Operand.src`) and the consuming
        `conv()` decides (in-kernel production, or `_materialize`)."""
        c1 = x1.C if x1 is not None else 0
        C = x0.C + c1
        shape = (self.B, x0.W * up + 2, x0.H * up, C)
        if not also_raw and self._emit(x0, x1, norm, silu, up, circular, terms):
            pr = x0.producer
            out = Operand(self.pg.alloc(shape, torch.float16, before_op=pr["index"]), None)
            op = pr["op"]
            op.p[19], op.p[20], op.p[21] = out.hi.data_ptr(), self.f32(norm.weight).data_ptr(), self.f32(norm.bias).data_ptr()
            op.i[20], op.i[21], op.i[22] = norm.num_groups, int(silu), int(circular)
            op.f[1] = norm.eps
            pr["emitted"] = True
            return out
        out = self.alloc_half(shape, terms)
        raw = self.alloc_half(shape, raw_terms) if also_raw else None
        spec = dict(x0=x0, x1=x1, norm=norm, silu=silu, up=up, circular=circular)
        fused_moments = norm is None or (x0.stats is not None and (x1 is None or x1.stats is not None)
                                         and (C // norm.num_groups) % 2 == 0 and x0.C % 2 == 0)
        out.src = spec
        if raw is not None:
            raw.src = dict(x0=x0, x1=x1, norm=None, silu=False, up=up, circular=circular)
        if not (defer and FUSE_PREP and CONV_KIND == _lib.OP_CONV_TC and fused_moments):
            self._materialize(out, raw)
        return (out, raw) if also_raw else out

    def _materialize(self, out, raw=None):
        """Emit the rldm_prep launch that produces `out` (and, from the same read, the raw operand `raw` OF THE SAME
        SOURCE tensors)."""
        assert raw is None or raw.src is None or (raw.src["x0"] is out.src["x0"] and raw.src["x1"] is out.src["x1"])
        sp = out.src
        x0, x1, norm, up = sp["x0"], sp["x1"], sp["norm"], sp["up"]
        c1 = x1.C if x1 is not None else 0
        C = x0.C + c1
        sums = pairs0 = pairs1 = gamma = beta = None
        eps, G = 0.0, 0
        if norm is not None:
            G = norm.num_groups
            fused = (x0.stats is not None and (x1 is None or x1.stats is not None) and (C // G) % 2 == 0
                     and x0.C % 2 == 0)
            if fused:     # channel-pair moments from the producing conv epilogues (both sources of a skip concat)
                pairs0, pairs1 = x0.stats, (x1.stats if x1 is not None else None)
            else:
                sums = self.gn_stats(x0, x1, G)
            gamma, beta, eps = self.f32(norm.weight), self.f32(norm.bias), norm.eps
        with_raw = raw is not None and raw.src is not None
        self.pg.add(_lib.OP_PREP, i=(x0.C, c1, G, int(sp["silu"]), up, self.B, x0.W, x0.H, int(sp["circular"])), f=(eps,),
                    p=(x0.t, x1.t if x1 is not None else None, sums, gamma, beta, out.hi, out.lo,
                       raw.hi if with_raw else None, raw.lo if with_raw else None, pairs0, pairs1))
        out.src = None
        if with_raw:
            raw.src = None

    def _emit(self, x0, x1, norm, silu, up, circular, terms):
        """True when the convolution that wrote x0 can produce this operand in its epilogue (rldm_conv_tc_emit)."""
        pr = getattr(x0, "producer", None)
        if not (EMIT_PREP and not FUSE_LEVELS and CONV_KIND == _lib.OP_CONV_TC and pr is not None and not pr.get("emitted")
                and x1 is None and up == 1 and norm is not None and terms < 3):
            return False
        g = pr["geom"]
        return bool(_lib.lib().rldm_conv_tc_emittable(*g, norm.num_groups))

    def _fusable(self, W, H, cin, cout, ks, stride, pad_lo, sc_cin, has_residual):
        """True when this convolution runs on the small-layer kernel, which can produce its own operand."""
        return bool(FUSE_PREP and CONV_KIND == _lib.OP_CONV_TC and
                    _lib.lib().rldm_conv_tc_fusable(self.B, W, H, cin, cout, ks, stride, pad_lo, sc_cin, int(has_residual)))

    def conv(self, xh, W, H, conv=None, packed=None, cin=None, cout=None, ks=3, stride=1, pad_lo=1, circular=True,
             temb=None, residual=None, stats=False, shortcut=None, terms=3):
        """fp16 clp operand pair -> fp32 cl Act (B,W/stride,H/stride,Cout).  stats=True: the epilogue also
        accumulates the GroupNorm moments of the output (consumed by the next prep instead of a gn_stats pass).
        shortcut=(operand pair, 1x1 conv module): that convolution is folded into the K loop of this launch."""
        if conv is not None:
            wt, bias = self.pack_conv(conv, terms)
            cout, cin, ks = conv.out_channels, conv.in_channels, conv.kernel_size[0]
            stride, pad_lo = conv.stride[0], conv.padding[0]
            circular = bool(getattr(conv, "circular", False))
            if conv.padding_mode != "zeros" or conv.groups != 1 or conv.dilation[0] != 1:
                raise NotImplementedError("conv with groups/dilation/padding_mode is outside the reference path")
        else:
            wt, bias = packed
        # operands that have not been produced yet: in-kernel production on the small-layer kernel, else a prep launch
        sc_opnd = shortcut[0] if shortcut is not None else None
        sc_cin_q = shortcut[1].in_channels if shortcut is not None else 0
        main_src = sc_src = None
        if xh.src is not None or (sc_opnd is not None and sc_opnd.src is not None):
            if self._fusable(W, H, cin, cout, ks, stride, pad_lo, sc_cin_q, residual is not None):
                main_src, sc_src = xh.src, (sc_opnd.src if sc_opnd is not None else None)
                xh.src = None
                if sc_opnd is not None:
                    sc_opnd.src = None
            else:
                if xh.src is not None:
                    self._materialize(xh)
                if sc_opnd is not None and sc_opnd.src is not None:
                    self._materialize(sc_opnd)
        Wo, Ho = W // stride, H // stride
        out = self.pg.alloc((self.B, Wo, Ho, cout))
        temb_t, temb_stride = (None, 0)
        if temb is not None:
            temb_t, temb_stride = temb
        kind = CONV_KIND
        ints = [temb_stride, self.B, W, H, cin, cout, ks, stride, pad_lo, int(circular)]
        st = None
        sc_ptrs = (None, None, None)
        if kind == _lib.OP_CONV_TC:
            if stats and FUSE_STATS and Wo * Ho >= 64:
                st = self.stats_slot(cout // 2)          # channel-pair moments [B][Cout/2][2]
            ints += [0]   # split_k: auto
            sc_cin = 0
            if shortcut is not None:
                sc_x, sc_conv = shortcut
                sc_wt, _ = self.pack_conv(sc_conv, terms)
                assert sc_conv.kernel_size == (1, 1) and sc_conv.out_channels == cout and stride == 1
                sc_cin = sc_conv.in_channels
                sc_ptrs = (sc_x[0], sc_x[1], sc_wt)
                bias = self._cached(("bias_sum", id(conv.bias), id(sc_conv.bias)),
                                    lambda: (self.f32(conv.bias) + self.f32(sc_conv.bias)).contiguous())
            ints += [sc_cin, terms]
        else:
            assert shortcut is None
        assert (xh[1] is not None) == (terms == 3), "operand planes do not match the layer's precision"
        src_ptrs, feps = (), ()
        if main_src is not None or sc_src is not None:
            def act_t(a):
                return a.t if a is not None else None
            m = main_src or dict(x0=None, x1=None, norm=None, silu=False, up=1)
            norm = m["norm"]
            mx0, mx1 = m["x0"], m["x1"]
            sx0, sx1 = (sc_src["x0"], sc_src["x1"]) if sc_src is not None else (None, None)
            src_ptrs = (act_t(mx0), act_t(mx1), mx0.stats if norm is not None else None,
                        mx1.stats if (norm is not None and mx1 is not None) else None,
                        self.f32(norm.weight) if norm is not None else None, self.f32(norm.bias) if norm is not None else None,
                        act_t(sx0), act_t(sx1))
            ints += [mx0.C if mx0 is not None else 0, mx1.C if mx1 is not None else 0, norm.num_groups if norm is not None else 0,
                     int(m["silu"]), m["up"], sx0.C if sx0 is not None else 0, sx1.C if sx1 is not None else 0]
            feps = (norm.eps if norm is not None else 0.0,)
        op = self.pg.add(kind, i=ints, f=feps, p=(xh[0], wt, bias, temb_t, residual.t if residual is not None else None, out,
                                                  xh[1], st) + sc_ptrs + src_ptrs, launches=1)
        producer = None
        if kind == _lib.OP_CONV_TC and main_src is None and sc_src is None:
            producer = dict(op=op, index=len(self.pg.ops) - 1,
                            geom=(self.B, W, H, cin, cout, ks, stride, pad_lo, sc_cin, int(residual is not None)))
        return Act(out, self.B, Wo, Ho, cout, st, producer)

    # ---- blocks ----------------------------------------------------------------------------
    def resnet(self, rb, x0, x1=None, free_inputs=True):
        """ResnetBlock2D on the virtual concat (x0 | x1) (App. A.1; `model.py:342-362`)."""
        pg = self.pg
        circ = lambda conv: bool(getattr(conv, "circular", False))
        fold = rb.conv_shortcut is not None and CONV_KIND == _lib.OP_CONV_TC and FUSE_SHORTCUT
        t1 = t2 = ts = self.terms(x0.W)
        xr = None
        if rb.conv_shortcut is not None:
            a1, xr = self.prep(x0, x1, rb.norm1, silu=True, circular=circ(rb.conv1), also_raw=True, terms=t1, raw_terms=ts,
                               defer=True)
        else:
            a1 = self.prep(x0, x1, rb.norm1, silu=True, circular=circ(rb.conv1), terms=t1, defer=True)
        c1m = rb.conv1
        if a1.src is not None and not self._fusable(x0.W, x0.H, c1m.in_channels, c1m.out_channels, 3, 1, 1, 0, False):
            self._materialize(a1, xr)       # one launch for the operand and the shortcut's raw operand, as before
        temb = None
        if rb.time_emb_proj is not None and self.temb is not None:
            tt, T = self.temb
            off = self.temb_rows[id(rb)]
            temb = (tt.view(-1)[off:], T)
        h = self.conv(a1, x0.W, x0.H, rb.conv1, temb=temb, stats=True, terms=t1)
        self.free_half(a1)
        a2 = self.prep(h, None, rb.norm2, silu=True, circular=circ(rb.conv2), terms=t2, defer=True)
        if h.producer is not None and h.producer.get("emitted"):
            # conv1 wrote conv2's operand itself and nothing else reads h: no fp32 store, no channel-pair moments
            h.producer["op"].p[5] = None
            h.producer["op"].p[7] = None
        if fold:
            # the 1x1 conv_shortcut rides in conv2's K loop (extra K steps over the raw operand)
            out = self.conv(a2, x0.W, x0.H, rb.conv2, stats=True, shortcut=(xr, rb.conv_shortcut), terms=t2)
            self.free_half(xr)
        elif rb.conv_shortcut is not None:
            if xr.src is not None:
                self._materialize(xr)
            sc = self.conv(xr, x0.W, x0.H, rb.conv_shortcut, terms=ts)
            self.free_half(xr)
            out = self.conv(a2, x0.W, x0.H, rb.conv2, residual=sc, stats=True, terms=t2)
            pg.free(sc.t)
        else:
            assert x1 is None
            out = self.conv(a2, x0.W, x0.H, rb.conv2, residual=x0, stats=True, terms=t2)
        self.free_half(a2)
        pg.free(h.t)                # (after conv2: it may read h itself when it produces its own operand)
        pg.taps.append((rb, out))
        if free_inputs:
            pg.free(x0.t)
            if x1 is not None:
                pg.free(x1.t)
        return out

    def attention(self, at, x, free_input=True):
        """Attention block with residual (App. A.1): GN -> qkv GEMM -> SDPA(d=8) -> out GEMM + x."""
        pg = self.pg
        if at.dim_head != 8:
            raise NotImplementedError(f"attention head_dim {at.dim_head}: the sm_100a attention kernel implements the "
                                      "reference's attention_head_dim=8")
        C = x.C
        t = self.terms(x.W)
        a = self.prep(x, None, at.group_norm, silu=False, terms=t, defer=True)
        qkv = self.conv(a, x.W, x.H, packed=self.pack_linear([at.to_q, at.to_k, at.to_v], t), cin=C, cout=3 * C, ks=1,
                        pad_lo=0, terms=t)
        self.free_half(a)
        o = self.alloc_half((self.B, x.W + 2, x.H, C), t)
        pg.add(_lib.OP_ATTENTION, i=(self.B, x.W * x.H, C, x.H), p=(qkv.t, o[0], o[1]))
        pg.free(qkv.t)
        out = self.conv(o, x.W, x.H, packed=self.pack_linear([at.to_out[0]], t), cin=C, cout=C, ks=1, pad_lo=0,
                        residual=x, stats=True, terms=t)
        self.free_half(o)
        pg.taps.append((at, out))
        if free_input:
            pg.free(x.t)
        return out

    def downsample(self, ds, x, free_input=True):
        """Patched Downsample2D (`ldm/utils.py:107-116`): raw cast + stride-2 conv; padding=0 is the
        VAE-encoder asymmetric pad (pad_lo = 0)."""
        t = self.terms(x.W)
        xr = self.prep(x, None, None, circular=bool(getattr(ds.conv, "circular", False)), terms=t, defer=True)
        out = self.conv(xr, x.W, x.H, ds.conv, stats=True, terms=t)
        self.free_half(xr)
        self.pg.taps.append((ds, out))
        if free_input:
            self.pg.free(x.t)
        return out

    def pack_conv_up2(self, conv, terms):
        """3x3 weights of a convolution that follows a nearest-2x upsampling, combined per OUTPUT PHASE (a, b) into the 2x2
        convolution the low-resolution input sees (rldm_conv_tc_up2): [4 phases][planes][4 taps][Cout][Cin] fp16."""
        def make():
            w = conv.weight.detach().to(self.pg.device, torch.float32)       # (Cout, Cin, kW, kH)
            # along one axis: phase 0 reads (i-1, i) with weights (k0, k1 + k2); phase 1 reads (i, i+1) with (k0 + k1, k2)
            comb = [((0,), (1, 2)), ((0, 1), (2,))]
            phases = []
            for a in range(2):
                for b in range(2):
                    taps = []
                    for ti in range(2):
                        for tj in range(2):
                            acc = 0
                            for kw in comb[a][ti]:
                                for kh in comb[b][tj]:
                                    acc = acc + w[:, :, kw, kh]
                            taps.append(acc)
                    wt = torch.stack(taps, 0)                                   # [4 taps][Cout][Cin]
                    hi = wt.to(torch.float16)
                    phases.append(hi if terms < 2 else torch.cat([hi, (wt - hi.float()).to(torch.float16)], 0))
            return self.pg.hold(torch.stack(phases, 0).contiguous())
        wt = self._cached(("conv_up2", id(conv.weight), min(terms, 2)), make)
        b = self.f32(conv.bias) if conv.bias is not None else None
        return wt, b

    def upsample(self, us, x):
        """Upsample2D (`model.py:120-125`): nearest 2x folded into the cast, then 3x3 conv -- or, for the layers the
        role-swapped kernel takes (the decoder's), folded into the convolution itself: four 2x2 phase convolutions over
        the low-resolution operand (rldm_conv_tc_up2; 4/9 of the multiply-adds, a quarter of the operand bytes)."""
        t = self.terms(x.W * 2)
        conv = us.conv
        if (FOLD_UPSAMPLE and CONV_KIND == _lib.OP_CONV_TC and conv.kernel_size == (3, 3) and conv.stride == (1, 1)
                and conv.padding == (1, 1) and conv.padding_mode == "zeros" and conv.groups == 1
                and _lib.lib().rldm_conv_tc_up2_ok(self.B, x.W, x.H, conv.in_channels, conv.out_channels)):
            circ = bool(getattr(conv, "circular", False))
            xr = self.prep(x, None, None, up=1, circular=circ, terms=t)
            wt, bias = self.pack_conv_up2(conv, t)
            cout = conv.out_channels
            out = self.pg.alloc((self.B, x.W * 2, x.H * 2, cout))
            st = self.stats_slot(cout // 2) if FUSE_STATS else None
            self.pg.add(_lib.OP_CONV_UP2, i=(self.B, x.W, x.H, conv.in_channels, cout, int(circ), t),
                        p=(xr[0], xr[1], wt, bias, out, st), launches=4)
            self.free_half(xr)
            act = Act(out, self.B, x.W * 2, x.H * 2, cout, st)
            self.pg.taps.append((us, act))
            self.pg.free(x.t)
            return act
        xr = self.prep(x, None, None, up=2, circular=bool(getattr(us.conv, "circular", False)), terms=t, defer=True)
        out = self.conv(xr, x.W * 2, x.H * 2, us.conv, stats=True, terms=t)
        self.free_half(xr)
        self.pg.taps.append((us, out))
        self.pg.free(x.t)
        return out

    def conv_in(self, conv, x0, c0, x1, c1, W, H):
        wt = self._cached(("conv_in", id(conv.weight)), lambda: conv.weight.detach().to(
            self.pg.device, torch.float32).permute(2, 3, 1, 0).contiguous())           # [9][Cin][Cout]
        out = self.pg.alloc((self.B, W, H, conv.out_channels))
        assert conv.in_channels == c0 + c1 and conv.kernel_size == (3, 3) and conv.padding == (1, 1)
        st = self.stats_slot(conv.out_channels // 2) if (FUSE_STATS and (W * H) % 32 == 0) else None
        self.pg.add(_lib.OP_CONV_IN, i=(c0, c1, self.B, W, H, conv.out_channels, int(getattr(conv, "circular", False))),
                    p=(x0, x1, wt, self.f32(conv.bias), out, st))
        act = Act(out, self.B, W, H, conv.out_channels, st)
        self.pg.taps.append((conv, act))
        return act

    def conv_out(self, norm, conv, x, out_ref):
        """conv_norm_out + SiLU + conv_out: one fused launch straight from the fp32 residual stream
        (RLDM_FUSE_CONV_OUT=0: rldm_prep + rldm_conv_out over the fp16 operand)."""
        wt = self._cached(("conv_out", id(conv.weight)), lambda: conv.weight.detach().to(
            self.pg.device, torch.float32).permute(2, 3, 0, 1).contiguous())           # [9][Cout][Cin]
        assert conv.kernel_size == (3, 3) and conv.padding == (1, 1)
        circ = int(getattr(conv, "circular", False))
        if FUSE_CONV_OUT and conv.out_channels in (2, 4, 8) and x.C % 4 == 0:
            G = norm.num_groups
            sums = pairs = None
            if x.stats is not None and (x.C // G) % 2 == 0:
                pairs = x.stats
            else:
                sums = self.gn_stats(x, None, G)
            self.pg.add(_lib.OP_NORM_CONV_OUT, i=(G, 1, self.B, x.W, x.H, x.C, conv.out_channels, circ), f=(norm.eps,),
                        p=(x.t, sums, pairs, self.f32(norm.weight), self.f32(norm.bias), wt, self.f32(conv.bias), out_ref))
        else:
            a = self.prep(x, None, norm, silu=True, circular=bool(circ), terms=3)
            self.pg.add(_lib.OP_CONV_OUT, i=(self.B, x.W, x.H, x.C, conv.out_channels, circ),
                        p=(a[0], wt, self.f32(conv.bias), out_ref, a[1]))
            self.free_half(a)
        self.pg.free(x.t)


def _timestep_to_float(timestep, B, device):
    if torch.is_tensor(timestep):
        t = timestep.detach().to(device=device, dtype=torch.float32).reshape(-1)
        return t.expand(B) if t.numel() == 1 else t
    return torch.full((B,), float(timestep), dtype=torch.float32, device=device)


class UNetPlan:
    """Compiled `UNet2DModel.forward` (App. A.1) for a fixed (batch, W, H).

    `x_in`  : (B, in_channels - cond_channels, W, H) fp32 ref layout (static input buffer)
    `cond`  : (B, cond_channels, W, H) or None -- second conv_in source (pos-encoding / condition),
              so samplers never materialise torch.cat([latents, cond], 1)
    `t_buf` : (B,) fp32 timesteps;  `out`: (B, out_channels, W, H) fp32 ref layout."""

    def __init__(self, model, batch, W, H, cond_channels=0, sampler=False):
        """sampler=True: the plan is one step of a fused trajectory (precision profile of the sampler, see PRECISION)."""
        dev = model.device
        _require_cuda_device(dev, "UNet2DModel")
        cfg = model.config
        self.B, self.W, self.H = batch, W, H
        L = len(cfg.block_out_channels)
        if W % (1 << (L - 1)) or H % (1 << (L - 1)):
            raise ValueError(f"sample size {(W, H)} must be divisible by {1 << (L - 1)}")
        pg = self.prog = Program(dev)
        bd = Builder(pg, batch, groups=cfg.norm_num_groups, cache=model._packed,
                     terms_of=lambda w: (PRECISION_TOP_SAMPLER if sampler else PRECISION_TOP) if w >= W else PRECISION)
        cin = cfg.in_channels
        self.x_in = pg.hold(torch.zeros(batch, cin - cond_channels, W, H, device=dev))
        self.cond = pg.hold(torch.zeros(batch, cond_channels, W, H, device=dev)) if cond_channels else None
        self.t_buf = pg.hold(torch.zeros(batch, device=dev))
        self.out = pg.hold(torch.zeros(batch, cfg.out_channels, W, H, device=dev))

        # ---- time embedding: MLP + every resnet projection in one stacked matrix ----
        resnets = [m for m in model.modules() if hasattr(m, "time_emb_proj") and m.time_emb_proj is not None]
        D0 = cfg.block_out_channels[0]
        D4 = model.time_embedding.linear_1.out_features
        rows, off = [], 0
        for r in resnets:
            bd.temb_rows[id(r)] = off
            rows.append(r.time_emb_proj)
            off += r.time_emb_proj.out_features
        T = off
        wp = pg.hold(torch.cat([r.weight.detach().float() for r in rows], 0).to(dev).contiguous())
        bp = pg.hold(torch.cat([r.bias.detach().float() for r in rows], 0).to(dev).contiguous())
        te = model.time_embedding
        scratch = pg.hold(torch.zeros(2, batch, D4, device=dev))
        temb_out = pg.hold(torch.zeros(batch, T, device=dev))
        pg.add(_lib.OP_TEMB, i=(batch, D0, D4, T),
               p=(self.t_buf, bd.f32(te.linear_1.weight), bd.f32(te.linear_1.bias), bd.f32(te.linear_2.weight),
                  bd.f32(te.linear_2.bias), wp, bp, scratch, temb_out), launches=3)
        self.temb_dims = (D0, D4)
        bd.temb = (temb_out, T)
        self.temb_out, self.temb_T = temb_out, T

        # ---- network ----
        h = bd.conv_in(model.conv_in, self.x_in, cin - cond_channels, self.cond, cond_channels, W, H)
        skips = [h]
        for blk in model.down_blocks:
            for j, rb in enumerate(blk.resnets):
                h = bd.resnet(rb, h, free_inputs=False)          # input is a skip (or feeds attention below)
                if getattr(blk, "attentions", None) is not None and not _is_identity_attn(blk.attentions[j]):
                    h = bd.attention(blk.attentions[j], h, free_input=True)
                skips.append(h)
            if blk.downsamplers is not None:
                h = bd.downsample(blk.downsamplers[0], h, free_input=False)
                skips.append(h)
        mb = model.mid_block
        h = bd.resnet(mb.resnets[0], h, free_inputs=False)       # h is still on the skip stack
        if not _is_identity_attn(mb.attentions[0]):
            h = bd.attention(mb.attentions[0], h)
        h = bd.resnet(mb.resnets[1], h)
        for blk in model.up_blocks:
            for j, rb in enumerate(blk.resnets):
                h = bd.resnet(rb, h, skips.pop())                # frees both inputs
                if getattr(blk, "attentions", None) is not None and not _is_identity_attn(blk.attentions[j]):
                    h = bd.attention(blk.attentions[j], h)
            if blk.upsamplers is not None:
                h = bd.upsample(blk.upsamplers[0], h)
        assert not skips
        bd.conv_out(model.conv_norm_out, model.conv_out, h, self.out)
        bd.finish()
        pg.finalize()

    def run(self, sample, timestep):
        if self.cond is not None:
            c0 = self.x_in.shape[1]
            self.x_in.copy_(sample[:, :c0])
            self.cond.copy_(sample[:, c0:])
        else:
            self.x_in.copy_(sample)
        self.t_buf.copy_(_timestep_to_float(timestep, self.B, self.t_buf.device))
        self.prog.run()
        return self.out.clone()


def _check_vae_identity(m, name):
    if not isinstance(m, nn.Identity):
        raise NotImplementedError(f"AutoencoderKL.{name} must be nn.Identity (as `ldm/inference.py:90-92` sets it for "
                                  "the reference checkpoints); a learned 1x1 quant conv is not implemented")


class VaeDecoderPlan:
    """Compiled `AutoencoderKL.decode` == sgm `Decoder.forward` (`model.py:1024-1057`)."""

    def __init__(self, vae, batch, W, H):
        dev = vae.device
        _require_cuda_device(dev, "AutoencoderKL")
        _check_vae_identity(vae.post_quant_conv, "post_quant_conv")
        dec = vae.decoder
        self.B = batch
        pg = self.prog = Program(dev)
        bd = Builder(pg, batch, groups=vae.config.norm_num_groups, cache=vae._packed,
                     terms_of=lambda w: PRECISION_VAE)
        zc = vae.config.latent_channels
        n_up = sum(1 for b in dec.up_blocks if b.upsamplers is not None)
        self.z_in = pg.hold(torch.zeros(batch, zc, W, H, device=dev))
        self.out = pg.hold(torch.zeros(batch, vae.config.out_channels, W << n_up, H << n_up, device=dev))
        h = bd.conv_in(dec.conv_in, self.z_in, zc, None, 0, W, H)
        mb = dec.mid_block
        h = bd.resnet(mb.resnets[0], h)
        if not _is_identity_attn(mb.attentions[0]):
            h = bd.attention(mb.attentions[0], h)
        h = bd.resnet(mb.resnets[1], h)
        for blk in dec.up_blocks:
            for rb in blk.resnets:
                h = bd.resnet(rb, h)
            if blk.upsamplers is not None:
                h = bd.upsample(blk.upsamplers[0], h)
        bd.conv_out(dec.conv_norm_out, dec.conv_out, h, self.out)
        bd.finish()
        pg.finalize()

    def run(self, z):
        self.z_in.copy_(z)
        self.prog.run()
        return self.out.clone()


class VaeEncoderPlan:
    """Compiled `AutoencoderKL.encode` moments == sgm `Encoder.forward` (`model.py:852-896`)."""

    def __init__(self, vae, batch, W, H):
        dev = vae.device
        _require_cuda_device(dev, "AutoencoderKL")
        _check_vae_identity(vae.quant_conv, "quant_conv")
        enc = vae.encoder
        self.B = batch
        pg = self.prog = Program(dev)
        n_down = sum(1 for b in enc.down_blocks if b.downsamplers is not None)
        bd = Builder(pg, batch, groups=vae.config.norm_num_groups, cache=vae._packed,
                     terms_of=lambda w: PRECISION_VAE)
        ic = vae.config.in_channels
        self.x_in = pg.hold(torch.zeros(batch, ic, W, H, device=dev))
        self.out = pg.hold(torch.zeros(batch, enc.conv_out.out_channels, W >> n_down, H >> n_down, device=dev))
        h = bd.conv_in(enc.conv_in, self.x_in, ic, None, 0, W, H)
        for blk in enc.down_blocks:
            for rb in blk.resnets:
                h = bd.resnet(rb, h)
            if blk.downsamplers is not None:
                h = bd.downsample(blk.downsamplers[0], h)
        mb = enc.mid_block
        h = bd.resnet(mb.resnets[0], h)
        if not _is_identity_attn(mb.attentions[0]):
            h = bd.attention(mb.attentions[0], h)
        h = bd.resnet(mb.resnets[1], h)
        bd.conv_out(enc.conv_norm_out, enc.conv_out, h, self.out)
        bd.finish()
        pg.finalize()

    def run(self, x):
        self.x_in.copy_(x)
        self.prog.run()
        return self.out.clone()
