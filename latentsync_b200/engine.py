"""Host-side execution engine: turns a state_dict into packed fp16 weights and a shape-specialised *plan* - a fixed
sequence of C-ABI kernel launches over pre-allocated HBM buffers - which is captured once into a CUDA graph and
replayed for every denoising step (the reference's eager forward is ~600 module calls).

Activation layout everywhere: channels-last fp16, a (b f) x H x W x C video tensor is the row-major matrix
[(b f h w), C].  Reference module semantics are cited at each builder function.

PyTorch is used for device memory, streams and CUDA-graph capture only; every arithmetic op is a kernel from
latentsync_b200/csrc reached through include/latentsync_b200.h.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .spec import SD_VAE_FT_MSE_CONFIG, unet_config, validate_unet_config

GEGLU_TILE = 256
KPAD = 64  # GEMM K granularity (one 128-byte swizzle row of fp16)
# nn.LayerNorm folded into the nn.Linear that consumes it (LsGemmArgs.col_sum / row_partials_*): the GEMM that writes the
# residual stream emits per-row (sum, sum of squares) partials from its epilogue, the next GEMM reads the RAW stream with
# W diag(gamma) and normalises in its own epilogue - no LayerNorm kernel, no normalised copy.  LS_FOLD_LN=0 keeps the
# explicit kernels (A/B measurements; the level with < FOLD_LN_MIN_ROWS rows always keeps them: its GEMMs are split-K).
FOLD_LN = os.environ.get("LS_FOLD_LN", "1") != "0"
FOLD_LN_MIN_ROWS = 2048
LN_TILE = 160  # tile width of every GEMM that emits partials (the cost model's own choice for N = 320 / 640 / 1280)
# nn.GroupNorm statistics from the epilogue of the GEMM that produces the normalised tensor (LsGemmArgs.gn_partials_out ->
# ls_groupnorm_parts): the norm becomes one read-modify-write pass without a reduction pass or a grid rendezvous.
# LS_GN_PARTS=0 keeps the self-contained kernels (A/B measurements).  Levels with < GN_PARTS_MIN_ROWS rows keep them too
# (their convolutions are split-K and the tensors are latency-, not bandwidth-bound).
GN_PARTS = os.environ.get("LS_GN_PARTS", "1") != "0"
GN_PARTS_MIN_ROWS = 2048
# Upsample3D / Upsample2D (nearest x2, then a 3x3 convolution) as four sub-pixel phase GEMMs over the LOW-resolution tensor
# (LsGemmArgs.up2, Plan.upconv): no upsampled copy and 4/9 of the multiply-adds.  LS_UPCONV_FOLD=0 keeps ls_upsample2x + the
# 3x3 convolution (A/B measurements); low-resolution tensors with < UPCONV_MIN_ROWS rows keep it too (four launches of
# M = 512 rows are slower than one convolution of 2048).
UPCONV_FOLD = os.environ.get("LS_UPCONV_FOLD", "1") != "0"
UPCONV_MIN_ROWS = 2048
# Downsample3D / Downsample2D (3x3 convolution, stride 2) read in place through TMA element strides (LsGemmArgs.stride2)
# instead of an explicit 9-tap im2col copy.  LS_S2_INPLACE=0 keeps ls_im2col_s2 + a plain GEMM (A/B measurements).
S2_INPLACE = os.environ.get("LS_S2_INPLACE", "1") != "0"


def ln_parts(n: int) -> int:
    """partials per row written by a producer GEMM of width n: three fixed column ranges per n-tile"""
    return 3 * ((n + LN_TILE - 1) // LN_TILE)


# --------------------------------------------------------------------------------------------------- buffers
class _Pool:
    """Exact-size free lists of device buffers.  A `Buf` returns its storage on garbage collection; because the plan is
    built in execution order and runs on one stream, a later op may safely overwrite a buffer whose last reader was
    recorded earlier.  Re-using storage keeps the working set inside the 126 MB L2 instead of streaming fresh lines."""

    def __init__(self, device):
        self.device = device
        self.free: Dict[int, List[torch.Tensor]] = {}
        self.all: List[torch.Tensor] = []

    def take(self, nbytes: int) -> torch.Tensor:
        lst = self.free.get(nbytes)
        if lst:
            return lst.pop()
        t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.all.append(t)
        return t

    def fresh(self, nbytes: int) -> torch.Tensor:
        """storage that no earlier launch of the plan has ever written (never taken from the free lists)"""
        t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.all.append(t)
        return t

    def give(self, t: torch.Tensor) -> None:
        self.free.setdefault(t.numel(), []).append(t)

    def total_bytes(self) -> int:
        return sum(t.numel() for t in self.all)


class Buf:
    """[rows, cols] device matrix handle; closures recorded in a plan capture only `.ptr` (an int)."""

    __slots__ = ("pool", "store", "ptr", "rows", "cols", "dtype", "aux", "gnp", "gnu")

    def __init__(self, pool: _Pool, rows: int, cols: int, dtype=torch.float16, fresh: bool = False):
        self.pool = pool
        self.aux = None  # fp32 [ln_parts(cols)][rows][2] LayerNorm partials written by the GEMM(s) that produced this buffer
        self.gnp = None  # fp32 [rows / 128][cols / gnu][2] GroupNorm partials written by the producing GEMM(s)
        self.gnu = 0     # columns per GroupNorm partial
        self.rows, self.cols, self.dtype = rows, cols, dtype
        nbytes = rows * cols * torch.empty((), dtype=dtype).element_size()
        nbytes = (nbytes + 255) // 256 * 256
        self.store = pool.fresh(nbytes) if fresh else pool.take(nbytes)
        self.ptr = self.store.data_ptr()

    def __del__(self):
        try:
            self.pool.give(self.store)
        except Exception:  # interpreter shutdown
            pass

    def tensor(self) -> torch.Tensor:
        n = self.rows * self.cols
        return self.store.view(self.dtype)[:n].view(self.rows, self.cols)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {L.lib().ls_last_error().decode()}")


class Plan:
    def __init__(self, device):
        self.device = torch.device(device)
        self.pool = _Pool(self.device)
        self.ops = []  # (callable, name)
        self.keep = []  # Bufs / tensors that must outlive the build (static inputs, outputs)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.lib = L.lib()
        self.stats_cursor = 0
        self.stats: Optional[torch.Tensor] = None
        self.launches = 0
        self.kinds: List[str] = []  # one tag per recorded launch (bench.py times kernels by kind)
        self.descs: List[str] = []
        self.op_flops: List[float] = []  # EXECUTED FLOPs of each launch (2*M*N*K for GEMM/conv, 4*B*h*Sq*Sk*d for SDPA)
        # ALGORITHMIC FLOPs of the reference operation the launch stands for: differs from the executed count only for the
        # sub-pixel phases of an upsample + 3x3 convolution (each is credited a quarter of the 9-tap convolution)
        self.op_flops_alg: List[float] = []
        self.op_bytes: List[float] = []  # algorithmic bytes of each GEMM launch: A once + W once + out (+ residual) once
        # indices of launches that do NOT depend on the loop state and are therefore kept out of the captured graph
        # (UNet: the time-embedding path = a function of t only, the audio K/V projection = a function of the audio only);
        # `run()` still runs everything in order, `replay()` runs the graph without them, `run_hoisted()` runs them alone
        self.hoisted: set = set()

    # ---- buffers
    def buf(self, rows, cols, dtype=torch.float16) -> Buf:
        return Buf(self.pool, rows, cols, dtype)

    def static(self, rows, cols, dtype=torch.float16) -> Buf:
        """a buffer that lives as long as the plan (inputs, outputs, zero-padded operands).  Always FRESH storage: a
        recycled block may still be the target of launches recorded earlier in the plan, which would overwrite what the
        builder puts here now (e.g. the zero padding columns of a narrow GEMM output) when the plan runs."""
        b = Buf(self.pool, rows, cols, dtype, fresh=True)
        self.keep.append(b)
        return b

    # ---- execution
    def run(self, hoisted: bool = True) -> None:
        for i, fn in enumerate(self.ops):
            if hoisted or i not in self.hoisted:
                fn()

    def run_hoisted(self, which=None) -> None:
        """the launches kept out of the graph (all of them, or the given index range), eagerly on the current stream"""
        for i in sorted(self.hoisted):
            if which is None or which[0] <= i < which[1]:
                self.ops[i]()

    def run_timed(self) -> Dict[str, Tuple[int, float]]:
        """eager run with a CUDA event pair around every launch: kind -> (launches, total ms).  Diagnostic only."""
        evs = []
        for fn in self.ops:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize(self.device)
        out: Dict[str, Tuple[int, float]] = {}
        self.last_times = [a.elapsed_time(b) for a, b in evs]
        for kind, ms1 in zip(self.kinds, self.last_times):
            n, ms = out.get(kind, (0, 0.0))
            out[kind] = (n + 1, ms + ms1)
        return out

    def shape_table(self, kind: str = "gemm"):
        """after run_timed(): [(desc, launches, total ms, TFLOP/s)] aggregated over identical launches"""
        agg: Dict[str, List[float]] = {}
        for k, d, f, ms in zip(self.kinds, self.descs, self.op_flops, self.last_times):
            if k == kind:
                e = agg.setdefault(d, [0, 0.0, 0.0])
                e[0] += 1
                e[1] += ms
                e[2] += f
        rows = [(d, int(n), ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0) for d, (n, ms, fl) in agg.items()]
        return sorted(rows, key=lambda r: -r[2])

    def capture(self) -> None:
        """warm up eagerly (sets function attributes, validates every launch), then record the CUDA graph"""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.run()
            self.run()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run(hoisted=False)
        self.graph = g

    def time_kind_in_graph(self, kind: Optional[str], reps: int = 5, exclude: Optional[str] = None) -> float:
        """ms per pass of ONLY the launches of one kind (e.g. "gemm"), captured in plan order into their own CUDA graph
        and timed with CUDA events: kernel time without the host launch gap that the per-launch event pairs of
        `run_timed` include (~3 us x launches).  Buffers keep whatever the last full run left in them (timing of these
        kernels does not depend on the values).  `kind=None, exclude="gemm"`: every launch EXCEPT that kind."""
        fns = [fn for i, (fn, k) in enumerate(zip(self.ops, self.kinds))
               if (kind is None or k == kind) and k != exclude and i not in self.hoisted]
        if not fns:
            return 0.0
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for fn in fns:
                fn()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for fn in fns:
                fn()
        g.replay()
        torch.cuda.synchronize(self.device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize(self.device)
        return a.elapsed_time(b) / reps

    def replay(self) -> None:
        """the graph (or, before capture, the same launches eagerly): everything except the hoisted launches"""
        if self.graph is None:
            self.run(hoisted=False)
        else:
            self.graph.replay()

    # ---- emitters (each records ONE launch)
    def _emit(self, fn, kind: str = "other", flops: float = 0.0, desc: str = "", nbytes: float = 0.0,
              alg_flops: Optional[float] = None) -> None:
        self.ops.append(fn)
        self.kinds.append(kind)
        self.op_flops.append(flops)
        self.op_flops_alg.append(flops if alg_flops is None else alg_flops)
        self.op_bytes.append(nbytes)
        self.descs.append(desc)
        self.launches += 1

    def flops(self, kind: Optional[str] = None, in_graph_only: bool = False, algorithmic: bool = False) -> float:
        # executed FLOPs of the plan's launches, or (algorithmic) the FLOPs of the reference operations they stand for
        return sum(f for i, (k, f) in enumerate(zip(self.kinds, self.op_flops_alg if algorithmic else self.op_flops))
                   if (kind is None or k == kind) and not (in_graph_only and i in self.hoisted))

    def bytes(self, kind: Optional[str] = None, in_graph_only: bool = False) -> float:
        return sum(b for i, (k, b) in enumerate(zip(self.kinds, self.op_bytes))
                   if (kind is None or k == kind) and not (in_graph_only and i in self.hoisted))

    def count(self, kind: str, in_graph_only: bool = False) -> int:
        return sum(1 for i, k in enumerate(self.kinds) if k == kind and not (in_graph_only and i in self.hoisted))

    def gemm(self, segs: Sequence[Tuple[int, int, int, int]], nimg: int, H: int, W: int, w: torch.Tensor, N: int,
             out_ptr: int, ldo: int, bias_ptr: int = 0, bias_div: int = 0, bias_ld: int = 0, residual_ptr: int = 0,
             ldr: int = 0, flags: int = 0, tile_n: int = 0, b_batch_stride: int = 0, b_ptr: int = 0,
             col_sum_ptr: int = 0, parts_in: Optional[Tuple[int, int, int]] = None,
             parts_out: Optional[Tuple[int, int, int]] = None, gn_out: Optional[Tuple[int, int, int]] = None,
             up2: int = 0, alg_flops: Optional[float] = None, stride2: int = 0, stride2_pad: int = 1) -> None:
        """segs: (ptr, channels, ld, taps).  w: packed fp16 [N, Ktot] tensor (kept alive by the engine) or b_ptr.
        parts_in / parts_out: (pointer to the first row, parts, stride in rows) of LayerNorm partials (LsGemmArgs).
        gn_out: (pointer to the first tile, unit, ld) of GroupNorm partials (LsGemmArgs.gn_partials_out)."""
        a = L.LsGemmArgs()
        a.up2 = up2  # 1 + 2 py + px: one sub-pixel phase of "nearest x2 upsample -> 3x3 conv" (segments with 4 taps)
        a.stride2, a.stride2_pad = stride2, stride2_pad  # stride-2 3x3 conv read in place: OUTPUT geometry, input tensor
        if gn_out is not None:
            a.gn_partials_out, a.gn_unit, a.gn_partials_ld = gn_out
        if parts_in is not None:
            a.row_partials_in, a.n_partials_in, a.partials_in_stride = parts_in
            a.col_sum, a.ln_eps = col_sum_ptr, 1e-5
        if parts_out is not None:
            a.row_partials_out, a.n_partials_out, a.partials_out_stride = parts_out
        a.nseg = len(segs)
        ktot = 0
        for i, (ptr, ch, ld, taps) in enumerate(segs):
            a.a_ptr[i], a.a_ch[i], a.a_ld[i], a.a_taps[i] = ptr, ch, ld, taps
            ktot += ch * taps
        a.nimg, a.H, a.W = nimg, H, W
        if w is not None:
            assert w.dtype == torch.float16 and w.is_contiguous() and w.shape[-1] == ktot and w.shape[0] == N, (
                tuple(w.shape), N, ktot)
            a.b_ptr = w.data_ptr()
        else:
            a.b_ptr = b_ptr
        a.N, a.Ktot, a.b_batch_stride = N, ktot, b_batch_stride
        a.bias = bias_ptr or None
        a.bias_div, a.bias_ld = bias_div, bias_ld
        a.residual = residual_ptr or None
        a.ldr = ldr
        a.out, a.ldo, a.flags, a.tile_n = out_ptr, ldo, flags, tile_n
        fn = self.lib.ls_gemm
        taps = "+".join(f"{ch}x{tp}" for _, ch, _, tp in segs)
        M = nimg * H * W
        n_out = N // 2 if flags & L.EPI_GEGLU else N
        nb = nimg if b_batch_stride else 1
        nbytes = (sum(M * ch * 2 for _, ch, _, _ in segs) + nb * N * ktot * 2 + M * n_out * (4 if flags & L.EPI_OUT_F32 else 2)
                  + (M * n_out * 2 if residual_ptr else 0))
        self._emit(lambda: _chk(fn(C.byref(a), _stream()), "ls_gemm"), "gemm", 2.0 * M * N * ktot,
                   f"M={M} N={N} K={ktot} img={nimg}x{H}x{W} segs={taps} flags={flags} batched={int(b_batch_stride != 0)}"
                   + (" res" if residual_ptr else "") + (" ln=in" if parts_in is not None else "")
                   + (" ln=out" if parts_out is not None else "") + (" gn=out" if gn_out is not None else "")
                   + (f" up2={up2}" if up2 else "") + (" stride2" if stride2 else ""),
                   float(nbytes), alg_flops)

    def upconv(self, x_ptr: int, cin: int, nimg: int, h: int, wd: int, phases: Sequence[torch.Tensor], cout: int,
               out_ptr: int, bias_ptr: int) -> None:
        """nearest x2 upsample + 3x3 convolution of a [nimg, h, wd, cin] tensor into [nimg, 2h, 2wd, cout]: four GEMM
        launches over the LOW-resolution tensor (pack_upconv_phases), each writing its quarter of the output pixels -
        no upsampled copy, 4/9 of the multiply-adds (resnet.py:47-75; diffusers Upsample2D)"""
        alg = 2.0 * (4 * nimg * h * wd) * cout * (9 * cin) / 4  # a quarter of the 9-tap convolution over the upsampled image
        for ph in range(4):
            self.gemm([(x_ptr, cin, cin, 4)], nimg, h, wd, phases[ph], cout, out_ptr, cout, bias_ptr=bias_ptr, up2=ph + 1,
                      alg_flops=alg)

    def begin_stats(self, nfloats: int) -> None:
        """one fp32 arena for the (sum, sumsq) results of every GroupNorm of the plan"""
        self.stats = torch.zeros(nfloats, dtype=torch.float32, device=self.device)

    def groupnorm(self, x1: int, c1: int, x2: int, c2: int, rows: int, rows_per_inst: int, groups: int,
                  gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool, out_ptr: int) -> None:
        ninst = rows // rows_per_inst
        n = ninst * groups * 2
        assert self.stats is not None and self.stats_cursor + n <= self.stats.numel(), "GroupNorm stats arena too small"
        sp = self.stats.data_ptr() + self.stats_cursor * 4
        self.stats_cursor += n
        fn = self.lib.ls_groupnorm
        g, b = gamma.data_ptr(), beta.data_ptr()
        x2p = x2 or None
        cc = c1 + c2
        self._emit(lambda: _chk(fn(x1, c1, x2p, c2, rows, rows_per_inst, groups, g, b, eps, int(silu), sp, out_ptr,
                                   _stream()), "ls_groupnorm"), "groupnorm", 0.0,
                   f"rows={rows} C={cc} rows_per_inst={rows_per_inst} silu={int(silu)}", 2.0 * rows * cc * 2)

    def groupnorm_parts(self, x1: int, c1: int, p1: int, ld1: int, x2: int, c2: int, p2: int, ld2: int, rows: int,
                        rows_per_inst: int, groups: int, unit: int, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                        silu: bool, out_ptr: int) -> None:
        """GroupNorm from the partials of the producing GEMM(s) (ls_groupnorm_parts)"""
        fn = self.lib.ls_groupnorm_parts
        g, b = gamma.data_ptr(), beta.data_ptr()
        x2p, p2p = (x2 or None), (p2 or None)
        cc = c1 + c2
        self._emit(lambda: _chk(fn(x1, c1, p1, ld1, x2p, c2, p2p, ld2, rows, rows_per_inst, groups, unit, g, b, eps,
                                   int(silu), out_ptr, _stream()), "ls_groupnorm_parts"), "groupnorm", 0.0,
                   f"rows={rows} C={cc} rows_per_inst={rows_per_inst} silu={int(silu)} parts", 2.0 * rows * cc * 2)

    def gn(self, srcs, rows: int, rows_per_inst: int, groups: int, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
           silu: bool, out_ptr: int) -> None:
        """GroupNorm of the virtual concatenation of srcs = [(Buf, channels)] (one or two), rows starting at each
        buffer's row 0: from the producers' partials when every source carries them, else the self-contained kernel"""
        (x1, c1) = srcs[0]
        (x2, c2) = srcs[1] if len(srcs) > 1 else (None, 0)
        cg = (c1 + c2) // groups
        unit = x1.gnu
        if (all(b.gnp is not None and b.gnu == unit for b, _ in srcs) and rows_per_inst % 128 == 0 and unit > 0
                and (c1 + c2) % groups == 0 and cg % unit == 0 and c1 % unit == 0):
            self.groupnorm_parts(x1.ptr, c1, x1.gnp.ptr, x1.cols // unit, x2.ptr if x2 is not None else 0, c2,
                                 x2.gnp.ptr if x2 is not None else 0, (x2.cols // unit) if x2 is not None else 0, rows,
                                 rows_per_inst, groups, unit, gamma, beta, eps, silu, out_ptr)
        else:
            self.groupnorm(x1.ptr, c1, x2.ptr if x2 is not None else 0, c2, rows, rows_per_inst, groups, gamma, beta, eps,
                           silu, out_ptr)

    def with_gn_parts(self, out: Buf, unit: int) -> Tuple[int, int, int]:
        """attach a GroupNorm-partials array to `out`; returns the gn_out argument of the GEMM that writes all of it"""
        assert out.rows % 128 == 0 and out.cols % unit == 0
        out.gnp = self.buf(out.rows // 128, (out.cols // unit) * 2, torch.float32)
        out.gnu = unit
        return (out.gnp.ptr, unit, out.cols // unit)

    def layernorm(self, x: int, rows: int, Cc: int, gamma: torch.Tensor, beta: torch.Tensor, out_ptr: int,
                  pe: Optional[torch.Tensor] = None, rows_per_frame: int = 1, nframes: int = 1) -> None:
        fn = self.lib.ls_layernorm
        g, b = gamma.data_ptr(), beta.data_ptr()
        pp = pe.data_ptr() if pe is not None else None
        self._emit(lambda: _chk(fn(x, rows, Cc, g, b, 1e-5, pp, rows_per_frame, nframes, out_ptr, _stream()),
                                "ls_layernorm"), "layernorm", 0.0, f"rows={rows} C={Cc} pe={int(pe is not None)}",
                   2.0 * rows * Cc * 2)

    def attention(self, q: int, k: int, v: int, out: int, ldq: int, ldk: int, ldv: int, ldo: int, batch: int,
                  heads: int, head_dim: int, sq: int, skv: int, q_addr=None, kv_addr=None) -> None:
        a = L.LsAttnArgs()
        a.q, a.k, a.v, a.out = q, k, v, out
        a.ldq, a.ldk, a.ldv, a.ldo = ldq, ldk, ldv, ldo
        a.batch, a.heads, a.head_dim, a.sq, a.skv = batch, heads, head_dim, sq, skv
        a.q_inner, a.q_outer_stride, a.q_inner_stride, a.q_seq_stride = q_addr or (1, sq, 0, 1)
        a.kv_inner, a.kv_outer_stride, a.kv_inner_stride, a.kv_seq_stride = kv_addr or (1, skv, 0, 1)
        a.scale = float(head_dim) ** -0.5
        fn = self.lib.ls_attention
        self._emit(lambda: _chk(fn(C.byref(a), _stream()), "ls_attention"), "attention",
                   4.0 * batch * heads * sq * skv * head_dim, f"batch={batch} heads={heads} d={head_dim} sq={sq} skv={skv}")

    def call(self, name: str, *args) -> None:
        """any other C-ABI function whose last parameter is the stream"""
        fn = getattr(self.lib, name)
        self._emit(lambda: _chk(fn(*args, _stream()), name), name[3:])


# ----------------------------------------------------------------------------------------------- weight packing
def _h(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float16).contiguous()


def pack_conv3x3(w: torch.Tensor, splits: Optional[Sequence[int]] = None) -> torch.Tensor:
    """OIHW -> fp16 [N, sum_seg 9*pad64(c_seg)], K index = (segment, tap=ky*3+kx, channel)"""
    n, cin = w.shape[0], w.shape[1]
    splits = list(splits) if splits else [cin]
    assert sum(splits) == cin
    parts, c0 = [], 0
    for cs in splits:
        cp = (cs + KPAD - 1) // KPAD * KPAD
        p = torch.zeros(n, w.shape[2], w.shape[3], cp, dtype=torch.float32, device=w.device)
        p[..., :cs] = w[:, c0:c0 + cs].permute(0, 2, 3, 1)
        parts.append(p.reshape(n, -1))
        c0 += cs
    return _h(torch.cat(parts, dim=1))


def pack_upconv_phases(w: torch.Tensor) -> List[torch.Tensor]:
    """OIHW 3x3 weight of a convolution that follows a nearest x2 upsample -> four fp16 [N, 4 * pad64(C)] matrices,
    index 2 py + px, K index = (a, b, channel) over the 2 x 2 low-resolution window of output pixel (2y + py, 2x + px)
    (LsGemmArgs.up2).  Window row a = 0 is low-resolution row y - 1 + py: with py = 0 only the tap ky = 0 lands on it and
    ky = 1, 2 both land on row y; with py = 1 ky = 0, 1 land on row y and ky = 2 on row y + 1 - likewise for columns.  Taps
    that land on the same source pixel are summed in fp32 before the rounding to fp16."""
    n, c = w.shape[0], w.shape[1]
    cp = (c + KPAD - 1) // KPAD * KPAD
    taps = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    out = []
    for py in (0, 1):
        for px in (0, 1):
            p = torch.zeros(n, 2, 2, cp, dtype=torch.float32, device=w.device)
            for a in (0, 1):
                for b in (0, 1):
                    p[:, a, b, :c] = w[:, :, taps[py][a], :][:, :, :, taps[px][b]].float().sum(dim=(2, 3))
            out.append(_h(p.reshape(n, 4 * cp)))
    return out


def pack_1x1(w: torch.Tensor) -> torch.Tensor:
    n, cin = w.shape[0], w.shape[1]
    cp = (cin + KPAD - 1) // KPAD * KPAD
    p = torch.zeros(n, cp, dtype=torch.float32, device=w.device)
    p[:, :cin] = w.reshape(n, cin)
    return _h(p)


class _Weights:
    """name -> packed device tensor; everything a plan points at lives here for the life of the engine"""

    def __init__(self, sd: Dict[str, torch.Tensor], device):
        self.sd = sd
        self.device = torch.device(device)
        self.t: Dict[str, torch.Tensor] = {}

    def raw(self, key: str) -> torch.Tensor:
        return self.sd[key].detach().to(self.device, torch.float32)

    def has(self, key: str) -> bool:
        return key in self.sd

    def put(self, name: str, t: torch.Tensor) -> torch.Tensor:
        self.t[name] = t.contiguous()
        return self.t[name]

    def f32(self, key: str) -> torch.Tensor:
        if key not in self.t:
            self.put(key, self.raw(key))
        return self.t[key]

    def lin(self, key: str) -> torch.Tensor:
        """Linear / 1x1-conv weight -> fp16 [N, K]"""
        name = key + "#h"
        if name not in self.t:
            self.put(name, pack_1x1(self.raw(key)))
        return self.t[name]

    def conv(self, key: str, splits=None) -> torch.Tensor:
        name = key + "#c"
        if name not in self.t:
            self.put(name, pack_conv3x3(self.raw(key), splits))
        return self.t[name]

    def upconv(self, key: str) -> List[torch.Tensor]:
        """the four sub-pixel phase matrices of a 3x3 convolution that follows a nearest x2 upsample"""
        names = [f"{key}#up{ph}" for ph in range(4)]
        if names[0] not in self.t:
            for nm, t in zip(names, pack_upconv_phases(self.raw(key))):
                self.put(nm, t)
        return [self.t[nm] for nm in names]

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.t.values())


# ------------------------------------------------------------------------------------------------ UNet engine
class UNetEngine:
    """UNet3DConditionModel.forward (latentsync/models/unet.py:312-471) as a CUDA-graph plan per input shape."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: dict, device="cuda"):
        self.cfg = unet_config(cfg)
        validate_unet_config(self.cfg)
        c = self.cfg
        self.device = torch.device(device)
        self.w = _Weights(state_dict, self.device)
        # the plan assumes bias-free q / k / v projections (attention.py:230-232); the null-audio shortcut depends on it
        biased = [k for k in state_dict if k.endswith((".to_q.bias", ".to_k.bias", ".to_v.bias"))]
        if biased:
            raise NotImplementedError(f"attention projections with bias are not implemented: {biased[:3]}")
        self.plans: Dict[Tuple[int, int, int, int, int], "UNetPlan"] = {}
        self._pack()

    # reference key helpers
    def _resnet_names(self) -> List[str]:
        c = self.cfg
        names = []
        for i in range(len(c["down_block_types"])):
            names += [f"down_blocks.{i}.resnets.{j}" for j in range(c["layers_per_block"])]
        names += ["mid_block.resnets.0", "mid_block.resnets.1"]
        for i in range(len(c["up_block_types"])):
            names += [f"up_blocks.{i}.resnets.{j}" for j in range(c["layers_per_block"] + 1)]
        return names

    def _attn_names(self) -> List[str]:
        c = self.cfg
        names = []
        for i, t in enumerate(c["down_block_types"]):
            if t == "CrossAttnDownBlock3D":
                names += [f"down_blocks.{i}.attentions.{j}" for j in range(c["layers_per_block"])]
        names += ["mid_block.attentions.0"]
        for i, t in enumerate(c["up_block_types"]):
            if t == "CrossAttnUpBlock3D":
                names += [f"up_blocks.{i}.attentions.{j}" for j in range(c["layers_per_block"] + 1)]
        return names

    def _pack(self) -> None:
        """one-time repack of the checkpoint tensors into GEMM operand layouts (fp16, K-major)"""
        w, c = self.w, self.cfg
        boc = c["block_out_channels"]
        # time path (unet.py:95-98,376-382) and all ResnetBlock3D.time_emb_proj (resnet.py:152) as ONE matrix
        w.put("te1.w", _h(w.raw("time_embedding.linear_1.weight")))
        w.put("te2.w", _h(w.raw("time_embedding.linear_2.weight")))
        tw, tb, cb, self.tproj_off = [], [], [], {}
        off = 0
        for r in self._resnet_names():
            tw.append(w.raw(r + ".time_emb_proj.weight"))
            tb.append(w.raw(r + ".time_emb_proj.bias"))
            cb.append(w.raw(r + ".conv1.bias"))
            self.tproj_off[r] = off
            off += tw[-1].shape[0]
        self.tproj_total = off
        w.put("tproj.w", _h(torch.cat(tw)))
        w.put("tproj.b", torch.cat(tb))
        w.put("tproj.add", torch.cat(cb))
        # audio cross-attention K/V projections of every block as ONE matrix (attention.py:231-232)
        self.kv_off = {}
        if c["add_audio_layer"]:
            kw, off = [], 0
            for a in self._attn_names():
                t = a + ".transformer_blocks.0.attn2"
                k, v = w.raw(t + ".to_k.weight"), w.raw(t + ".to_v.weight")
                self.kv_off[a] = (off, off + k.shape[0])
                off += k.shape[0] + v.shape[0]
                kw += [k, v]
            self.kv_total = off
            kcat = torch.cat(kw)
            w.put("kv.w", pack_1x1(kcat))
        self.cross_k = (c["cross_attention_dim"] + KPAD - 1) // KPAD * KPAD

    def time_table(self, timesteps) -> torch.Tensor:
        """[len(timesteps), tproj_total] fp32: for every timestep of a denoising loop, each ResnetBlock3D's
        time_emb_proj(SiLU(time_embedding(t))) + conv1.bias (unet.py:376-382, resnet.py:190-205) - the same four launches
        the plan makes for one t, batched over the loop's timesteps (they depend on t only, lipsync_pipeline.py:537-554)."""
        lib, w, c = L.lib(), self.w, self.cfg
        boc = c["block_out_channels"]
        T = len(timesteps)
        t = torch.tensor([float(v) for v in timesteps], dtype=torch.float32, device=self.device)
        te0 = torch.empty(T, boc[0], dtype=torch.float32, device=self.device)
        te1 = torch.empty(T, boc[0] * 4, dtype=torch.float32, device=self.device)
        emb = torch.empty_like(te1)
        table = torch.empty(T, self.tproj_total, dtype=torch.float32, device=self.device)
        st = _stream()
        _chk(lib.ls_timestep_embedding(t.data_ptr(), T, boc[0], te0.data_ptr(), st), "ls_timestep_embedding")
        _chk(lib.ls_small_linear(te0.data_ptr(), T, boc[0], w.t["te1.w"].data_ptr(),
                                 w.f32("time_embedding.linear_1.bias").data_ptr(), None, boc[0] * 4, 0, 1,
                                 te1.data_ptr(), st), "ls_small_linear")
        _chk(lib.ls_small_linear(te1.data_ptr(), T, boc[0] * 4, w.t["te2.w"].data_ptr(),
                                 w.f32("time_embedding.linear_2.bias").data_ptr(), None, boc[0] * 4, 0, 0,
                                 emb.data_ptr(), st), "ls_small_linear")
        _chk(lib.ls_small_linear(emb.data_ptr(), T, boc[0] * 4, w.t["tproj.w"].data_ptr(), w.t["tproj.b"].data_ptr(),
                                 w.t["tproj.add"].data_ptr(), self.tproj_total, 1, 0, table.data_ptr(), st),
             "ls_small_linear")
        return table

    # ---- packed-weight accessors used by the plan builder
    def qkv(self, p: str) -> torch.Tensor:
        name = p + "#qkv"
        if name not in self.w.t:
            self.w.put(name, _h(torch.cat([self.w.raw(p + ".to_q.weight"), self.w.raw(p + ".to_k.weight"),
                                            self.w.raw(p + ".to_v.weight")])))
        return self.w.t[name]

    def folded(self, name: str, ln: str, weight_fn, bias_key: Optional[str] = None, geglu: bool = False,
               pe: Optional[torch.Tensor] = None, repeat: int = 1):
        """(W diag(gamma) fp16, col_sum fp32, bias' fp32) for LayerNorm `ln` followed by the Linear whose [N, K] weight
        `weight_fn()` returns (_lib.fold_layernorm).  GEGLU projections are re-ordered per N tile like `geglu()`; with
        `pe` ([F, K]) bias' is the per-frame table [(repeat F), N] (one row per (batch element, frame))."""
        key = f"{name}#ln" + (f"pe{pe.shape[0]}x{repeat}" if pe is not None else "")
        w = self.w
        if key + ".w" not in w.t:
            bias = w.raw(bias_key) if bias_key else None
            wg, cs, b2 = L.fold_layernorm(weight_fn(), bias, w.raw(ln + ".weight"), w.raw(ln + ".bias"), pe=pe)
            if geglu:
                wg, b2 = L.pack_geglu(wg, b2, GEGLU_TILE)
                cs = L.pack_geglu(cs[:, None], None, GEGLU_TILE)[0][:, 0]
            if pe is not None and repeat > 1:
                b2 = b2.repeat(repeat, 1)
            w.put(key + ".w", wg.to(torch.float16))
            w.put(key + ".cs", cs.float())
            w.put(key + ".b", b2.float())
        return w.t[key + ".w"], w.t[key + ".cs"], w.t[key + ".b"]

    def geglu(self, p: str) -> Tuple[torch.Tensor, torch.Tensor]:
        name = p + "#geglu"
        if name + ".w" not in self.w.t:
            wp, bp = L.pack_geglu(self.w.raw(p + ".weight"), self.w.raw(p + ".bias"), GEGLU_TILE)
            self.w.put(name + ".w", _h(wp))
            self.w.put(name + ".b", bp)
        return self.w.t[name + ".w"], self.w.t[name + ".b"]

    def conv2_with_shortcut(self, r: str, splits: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """conv2 (3x3) and the 1x1 conv_shortcut of a ResnetBlock3D fused into one GEMM along K (resnet.py:215-221)"""
        name = r + "#c2sc"
        if name + ".w" not in self.w.t:
            w2 = pack_conv3x3(self.w.raw(r + ".conv2.weight"))
            ws = self.w.raw(r + ".conv_shortcut.weight")
            parts, c0 = [w2], 0
            for cs in splits:
                parts.append(pack_1x1(ws[:, c0:c0 + cs]))
                c0 += cs
            self.w.put(name + ".w", torch.cat(parts, dim=1))
            self.w.put(name + ".b", self.w.raw(r + ".conv2.bias") + self.w.raw(r + ".conv_shortcut.bias"))
        return self.w.t[name + ".w"], self.w.t[name + ".b"]

    def plan(self, B: int, F: int, H: int, W: int, S: int, uncond_zero: bool = False,
             same_sample: bool = False) -> "UNetPlan":
        """`uncond_zero`: the caller guarantees that the FIRST half of the batch carries all-zero audio embeddings (the
        pipeline builds the classifier-free-guidance batch that way, lipsync_pipeline.py:503-507).  With bias-free
        to_k / to_v (attention.py:231-232) its keys and values are exactly 0, softmax is uniform over zeros, the
        attention output is exactly 0 and `to_out` adds exactly its bias: those rows skip LayerNorm, to_q, attention and
        to_out and get `hidden + bias` instead - the same bits, fewer launches' worth of rows (B == 2 only)."""
        uncond_zero = bool(uncond_zero) and B == 2 and S > 0
        # `same_sample` (only together with uncond_zero): both halves of the batch also carry the SAME sample and timestep
        # (lipsync_pipeline.py:542-549 duplicates the latents).  Until the first audio cross-attention the two halves
        # are then the same tensor: the first ResnetBlock3D and the self-attention part of the first transformer run
        # once, on half the rows.
        same_sample = bool(same_sample) and uncond_zero
        key = (B, F, H, W, S, uncond_zero, same_sample)
        if key not in self.plans:
            self.plans[key] = UNetPlan(self, B, F, H, W, S, uncond_zero, same_sample)
        return self.plans[key]


class UNetPlan(Plan):
    def __init__(self, eng: UNetEngine, B: int, F: int, H: int, W: int, S: int, uncond_zero: bool = False,
                 same_sample: bool = False):
        super().__init__(eng.device)
        self.eng = eng
        self.B, self.F, self.H, self.W, self.S = B, F, H, W, S
        self._full_B = B  # self.B is temporarily 1 inside the shared prefix of same_sample plans
        self.uncond_zero = uncond_zero
        self.same_sample = same_sample
        self.taps: Dict[str, Tuple[Buf, int]] = {}  # name -> (buffer, level); filled only when LS_DEBUG_TAPS=1
        self.debug = os.environ.get("LS_DEBUG_TAPS", "0") == "1"
        c = eng.cfg
        nlev = len(c["block_out_channels"])
        assert H % (1 << (nlev - 1)) == 0 and W % (1 << (nlev - 1)) == 0, "latent size must divide by 2^(levels-1)"
        self._build()

    # -- geometry of level l
    def _geo(self, lvl: int):
        return self.H >> lvl, self.W >> lvl

    def _rows(self, lvl: int) -> int:
        h, w = self._geo(lvl)
        return self.B * self.F * h * w

    def _tap(self, name: str, x: Buf, lvl: int) -> None:
        if self.debug:
            self.taps[name] = (x, lvl)

    def _gn_unit(self, lvl: int, n: int) -> int:
        """columns per GroupNorm partial for a block-level tensor of level `lvl` and width `n` whose producing GEMM emits
        the statistics of its output, or 0: that level keeps the self-contained GroupNorm kernels.  Like `_fold` the
        decision depends on the level's full-batch geometry only, so that plan variants stay bit-identical."""
        c = self.eng.cfg
        unit = c["block_out_channels"][0] // c["norm_num_groups"]  # divides C / groups of every norm, concatenations included
        h, wd = self._geo(lvl)
        per_b = self.F * h * wd
        if (not GN_PARTS or self._full_B * per_b < GN_PARTS_MIN_ROWS or per_b % 128 != 0 or unit < 1 or n % unit != 0
                or any(ch % unit for ch in c["block_out_channels"])):
            return 0
        return unit if any(n % bn == 0 and bn % unit == 0 for bn in range(32, 257, 32)) else 0

    def tap_tensor(self, name: str) -> torch.Tensor:
        """debug: activation `name` as (B, C, F, h, w) fp32"""
        x, lvl = self.taps[name]
        h, w = self._geo(lvl)
        return x.tensor().float().reshape(self.B, self.F, h, w, x.cols).permute(0, 4, 1, 2, 3).contiguous()

    # -- module builders --------------------------------------------------------------------------------------
    def _resnet(self, r: str, srcs: List[Tuple[Buf, int]], cout: int, lvl: int) -> Buf:
        """ResnetBlock3D.forward (resnet.py:182-223). srcs = [(hidden, c1)] or [(hidden, c1), (skip, c2)]: the
        torch.cat of unet_blocks.py:624,745 is never materialised for the convs; GroupNorm statistics span
        (C/32, F, H, W) of the virtual concatenation per batch element (plain nn.GroupNorm on a 5-D tensor)."""
        eng, w, c = self.eng, self.eng.w, self.eng.cfg
        h, wd = self._geo(lvl)
        rows = self._rows(lvl)
        per_b = self.F * h * wd
        cin = sum(ch for _, ch in srcs)
        x1, c1 = srcs[0]
        g, eps = c["norm_num_groups"], c["norm_eps"]
        gu = self._gn_unit(lvl, cout)  # both convolutions emit the GroupNorm statistics of what they store
        y1 = self.buf(rows, cin)
        self.gn(srcs, rows, per_b, g, w.f32(r + ".norm1.weight"), w.f32(r + ".norm1.bias"), eps, True, y1.ptr)
        h1 = self.buf(rows, cout)
        toff = eng.tproj_off[r]
        self.gemm([(y1.ptr, cin, cin, 9)], self.B * self.F, h, wd, w.conv(r + ".conv1.weight"), cout, h1.ptr, cout,
                  bias_ptr=self.tproj.ptr + toff * 4, bias_div=per_b, bias_ld=eng.tproj_total,
                  gn_out=self.with_gn_parts(h1, gu) if gu else None)
        del y1
        y2 = self.buf(rows, cout)
        self.gn([(h1, cout)], rows, per_b, g, w.f32(r + ".norm2.weight"), w.f32(r + ".norm2.bias"), eps, True, y2.ptr)
        del h1
        out = self.buf(rows, cout)
        gno = self.with_gn_parts(out, gu) if gu else None
        if w.has(r + ".conv_shortcut.weight"):
            wp, bp = eng.conv2_with_shortcut(r, [ch for _, ch in srcs])
            segs = [(y2.ptr, cout, cout, 9)] + [(b.ptr, ch, ch, 1) for b, ch in srcs]
            self.gemm(segs, self.B * self.F, h, wd, wp, cout, out.ptr, cout, bias_ptr=bp.data_ptr(), gn_out=gno)
        else:
            assert len(srcs) == 1 and c1 == cout
            self.gemm([(y2.ptr, cout, cout, 9)], self.B * self.F, h, wd, w.conv(r + ".conv2.weight"), cout, out.ptr,
                      cout, bias_ptr=w.f32(r + ".conv2.bias").data_ptr(), residual_ptr=x1.ptr, ldr=cout, gn_out=gno)
        return out

    def _fold(self, lvl: int, cc: int, pe: bool = False, halves: bool = False) -> bool:
        """whether a LayerNorm of level `lvl` (width `cc`) is folded into the Linear that consumes it.  The decision only
        depends on the level's full-batch geometry, never on the plan variant, so that the null-audio / shared-prefix
        plans stay bit-identical to the full plan: `halves` marks the norms around the audio cross-attention, whose GEMMs
        run on half the rows in those variants (they must still be >= FOLD_LN_MIN_ROWS: smaller GEMMs are split-K).  With
        the temporal sinusoid table (`pe`) the per-frame bias row must be constant over a 128-row tile."""
        h, wd = self._geo(lvl)
        rows = self._full_B * self.F * h * wd
        return (FOLD_LN and rows >= FOLD_LN_MIN_ROWS * (2 if halves else 1) and cc % 32 == 0
                and (not pe or (h * wd) % 128 == 0))

    def _with_parts(self, out: Buf) -> Buf:
        out.aux = self.buf(ln_parts(out.cols) * out.rows, 2, torch.float32)
        return out

    @staticmethod
    def _parts(x: Buf, row0: int = 0) -> Tuple[int, int, int]:
        """(pointer, parts, stride) of x's LayerNorm partials starting at row `row0`"""
        return (x.aux.ptr + row0 * 8, ln_parts(x.cols), x.rows)

    def _linear(self, x: Buf, key: str, n: int, bias: bool = True, residual: Optional[Buf] = None,
                stats: bool = False, gn: int = 0) -> Buf:
        """`stats`: the output feeds a LayerNorm that is folded into the next GEMM - this GEMM's epilogue emits the
        per-row partial sums of what it stores (out.aux).  `gn` (columns per partial): the output feeds a GroupNorm -
        the epilogue emits its per-tile statistics (out.gnp)."""
        w = self.eng.w
        k = w.lin(key + ".weight").shape[1]
        assert k == x.cols and not (stats and gn)
        out = self.buf(x.rows, n)
        if stats:
            self._with_parts(out)
        self.gemm([(x.ptr, k, k, 1)], 1, 1, x.rows, w.lin(key + ".weight"), n, out.ptr, n,
                  bias_ptr=w.f32(key + ".bias").data_ptr() if bias else 0,
                  residual_ptr=residual.ptr if residual is not None else 0, ldr=n,
                  tile_n=LN_TILE if stats else 0, parts_out=self._parts(out) if stats else None,
                  gn_out=self.with_gn_parts(out, gn) if gn else None)
        return out

    def _ff(self, ln: str, ff: str, hs: Buf, stats_next: bool = False) -> Buf:
        """x += FF(LN(x)); diffusers FeedForward/GEGLU (attention.py:171,197 ; motion_module.py:200,216)"""
        w = self.eng.w
        cc = hs.cols
        gg = self.buf(hs.rows, 4 * cc)
        if hs.aux is not None:  # LayerNorm folded into the GEGLU projection (the producer of hs emitted the partials)
            wp, cs, bp = self.eng.folded(ff + ".net.0.proj", ln, lambda: w.raw(ff + ".net.0.proj.weight"),
                                         ff + ".net.0.proj.bias", geglu=True)
            self.gemm([(hs.ptr, cc, cc, 1)], 1, 1, hs.rows, wp, 8 * cc, gg.ptr, 4 * cc, bias_ptr=bp.data_ptr(),
                      flags=L.EPI_GEGLU, tile_n=GEGLU_TILE, col_sum_ptr=cs.data_ptr(), parts_in=self._parts(hs))
        else:
            n = self.buf(hs.rows, cc)
            self.layernorm(hs.ptr, hs.rows, cc, w.f32(ln + ".weight"), w.f32(ln + ".bias"), n.ptr)
            wp, bp = self.eng.geglu(ff + ".net.0.proj")
            self.gemm([(n.ptr, cc, cc, 1)], 1, 1, hs.rows, wp, 8 * cc, gg.ptr, 4 * cc, bias_ptr=bp.data_ptr(),
                      flags=L.EPI_GEGLU, tile_n=GEGLU_TILE)
            del n
        return self._linear(gg, ff + ".net.2", cc, residual=hs, stats=stats_next)

    def _transformer(self, a: str, x: Buf, cc: int, lvl: int, shared: bool = False) -> Buf:
        """Transformer3DModel.forward (attention.py:82-124) + BasicTransformerBlock.forward (:174-199).
        `shared` (same_sample plans, first transformer only): `x` holds ONE batch element that stands for both; the part
        before the audio cross-attention runs on it once."""
        eng, w, c = self.eng, self.eng.w, self.eng.cfg
        h, wd = self._geo(lvl)
        full_B = self.B
        if shared:
            assert self.uncond_zero and full_B == 2 and x.rows * 2 == self._rows(lvl)
            self.B = 1  # geometry of the shared part: one batch element
        rows, hw = self._rows(lvl), h * wd
        heads = c["attention_head_dim"]
        d = cc // heads
        nrm = self.buf(rows, cc)
        self.gn([(x, cc)], rows, hw, c["norm_num_groups"], w.f32(a + ".norm.weight"), w.f32(a + ".norm.bias"), 1e-6,
                False, nrm.ptr)
        hs = self._linear(nrm, a + ".proj_in", cc, stats=self._fold(lvl, cc))
        del nrm
        t = a + ".transformer_blocks.0"
        # self-attention over the h*w tokens of each frame
        n = self.buf(rows, cc)
        qkv = self.buf(rows, 3 * cc)
        if hs.aux is not None:
            a1 = t + ".attn1"
            wq, cs, bq = eng.folded(a1 + ".qkv", t + ".norm1", lambda: torch.cat(
                [w.raw(a1 + ".to_q.weight"), w.raw(a1 + ".to_k.weight"), w.raw(a1 + ".to_v.weight")]))
            self.gemm([(hs.ptr, cc, cc, 1)], 1, 1, rows, wq, 3 * cc, qkv.ptr, 3 * cc, bias_ptr=bq.data_ptr(),
                      col_sum_ptr=cs.data_ptr(), parts_in=self._parts(hs))
        else:
            self.layernorm(hs.ptr, rows, cc, w.f32(t + ".norm1.weight"), w.f32(t + ".norm1.bias"), n.ptr)
            self.gemm([(n.ptr, cc, cc, 1)], 1, 1, rows, eng.qkv(t + ".attn1"), 3 * cc, qkv.ptr, 3 * cc)
        o = n  # reuse
        self.attention(qkv.ptr, qkv.ptr + 2 * cc, qkv.ptr + 4 * cc, o.ptr, 3 * cc, 3 * cc, 3 * cc, cc, self.B * self.F,
                       heads, d, hw, hw)
        del qkv
        cross = bool(c["add_audio_layer"] and self.audio_kv is not None)
        # the next LayerNorm is norm2 -> to_q (the conditional half only when the first half carries null audio), or
        # norm3 -> GEGLU projection when there is no audio layer
        hs2 = self._linear(o, t + ".attn1.to_out.0", cc, residual=hs, stats=self._fold(lvl, cc, halves=cross))
        del o, n, hs
        hs = hs2
        if shared:
            self.B = full_B
            rows = self._rows(lvl)
        if c["add_audio_layer"] and self.audio_kv is not None and self.uncond_zero:
            # cross-attention with a null-audio first half (see UNetEngine.plan): only the conditional rows go through
            # LayerNorm -> to_q -> attention -> to_out; the unconditional rows become hidden + to_out.bias
            half = rows // 2
            off = half * cc * 2  # bytes to the conditional half of an fp16 [rows][cc] buffer
            hoff = 0 if hs.rows == half else off  # a shared (one-element) hidden state serves both halves
            nh = self.buf(half, cc)
            qh = self.buf(half, cc)
            if hs.aux is not None:
                wq, cs, bq = eng.folded(t + ".attn2.to_q", t + ".norm2", lambda: w.raw(t + ".attn2.to_q.weight"))
                self.gemm([(hs.ptr + hoff, cc, cc, 1)], 1, 1, half, wq, cc, qh.ptr, cc, bias_ptr=bq.data_ptr(),
                          col_sum_ptr=cs.data_ptr(), parts_in=self._parts(hs, hoff // (cc * 2)))
            else:
                self.layernorm(hs.ptr + hoff, half, cc, w.f32(t + ".norm2.weight"), w.f32(t + ".norm2.bias"), nh.ptr)
                self.gemm([(nh.ptr, cc, cc, 1)], 1, 1, half, w.lin(t + ".attn2.to_q.weight"), cc, qh.ptr, cc)
            koff, voff = eng.kv_off[a]
            ldkv = eng.kv_total
            kv_rows = self.F * self.S * ldkv * 2  # bytes to the conditional half of audio_kv
            self.attention(qh.ptr, self.audio_kv.ptr + kv_rows + 2 * koff, self.audio_kv.ptr + kv_rows + 2 * voff,
                           nh.ptr, cc, ldkv, ldkv, cc, self.F, heads, d, hw, self.S)
            del qh
            hs2 = self.buf(rows, cc)
            st3 = self._fold(lvl, cc, halves=True)  # norm3 folded into the GEGLU projection: both halves' producers emit partials
            if st3:
                self._with_parts(hs2)
            tn = LN_TILE if st3 else 0
            ob = w.f32(t + ".attn2.to_out.0.bias").data_ptr()
            self.gemm([(nh.ptr, cc, cc, 1)], 1, 1, half, w.lin(t + ".attn2.to_out.0.weight"), cc, hs2.ptr + off, cc,
                      bias_ptr=ob, residual_ptr=hs.ptr + hoff, ldr=cc, tile_n=tn,
                      parts_out=self._parts(hs2, half) if st3 else None)
            # unconditional rows: a K = 64 GEMM over zero operands is `bias + residual` in the GEMM's own epilogue
            if getattr(self, "_zero_a", None) is None:
                self._zero_a = torch.zeros(self._rows(0) // 2, KPAD, dtype=torch.float16, device=eng.device)
                self._zero_w = torch.zeros(max(c["block_out_channels"]), KPAD, dtype=torch.float16, device=eng.device)
            self.gemm([(self._zero_a.data_ptr(), KPAD, KPAD, 1)], 1, 1, half, self._zero_w[:cc], cc, hs2.ptr, cc,
                      bias_ptr=ob, residual_ptr=hs.ptr, ldr=cc, tile_n=tn, parts_out=self._parts(hs2) if st3 else None)
            del nh, hs
            hs = hs2
        elif c["add_audio_layer"] and self.audio_kv is not None:
            # cross-attention: frame f attends to its own S audio tokens (attention.py:183-194)
            n = self.buf(rows, cc)
            if hs.aux is not None:
                wq, cs, bq = eng.folded(t + ".attn2.to_q", t + ".norm2", lambda: w.raw(t + ".attn2.to_q.weight"))
                q = self.buf(rows, cc)
                self.gemm([(hs.ptr, cc, cc, 1)], 1, 1, rows, wq, cc, q.ptr, cc, bias_ptr=bq.data_ptr(),
                          col_sum_ptr=cs.data_ptr(), parts_in=self._parts(hs))
            else:
                self.layernorm(hs.ptr, rows, cc, w.f32(t + ".norm2.weight"), w.f32(t + ".norm2.bias"), n.ptr)
                q = self._linear(n, t + ".attn2.to_q", cc, bias=False)
            koff, voff = eng.kv_off[a]
            ldkv = eng.kv_total
            self.attention(q.ptr, self.audio_kv.ptr + 2 * koff, self.audio_kv.ptr + 2 * voff, n.ptr, cc, ldkv, ldkv,
                           cc, self.B * self.F, heads, d, hw, self.S)
            del q
            hs2 = self._linear(n, t + ".attn2.to_out.0", cc, residual=hs, stats=self._fold(lvl, cc, halves=True))
            del n, hs
            hs = hs2
        hs = self._ff(t + ".norm3", t + ".ff", hs)
        if shared:
            # residual = the shared input for both halves: proj_out as two half-batch GEMMs
            key = a + ".proj_out"
            out = self.buf(rows, cc)
            half = rows // 2
            gu = self._gn_unit(lvl, cc)
            gno = self.with_gn_parts(out, gu) if gu else None
            for e in range(2):
                o8 = e * half * cc * 2
                g8 = (gno[0] + e * (half // 128) * gno[2] * 8, gno[1], gno[2]) if gno else None  # tiles of this half
                self.gemm([(hs.ptr + o8, cc, cc, 1)], 1, 1, half, w.lin(key + ".weight"), cc, out.ptr + o8, cc,
                          bias_ptr=w.f32(key + ".bias").data_ptr(), residual_ptr=x.ptr, ldr=cc, gn_out=g8)
            return out
        return self._linear(hs, a + ".proj_out", cc, residual=x, gn=self._gn_unit(lvl, cc))

    def _motion(self, m: str, x: Buf, cc: int, lvl: int) -> Buf:
        """VanillaTemporalModule (motion_module.py:126-151,203-218,262-313).  The "(b f) s c -> (b s) f c" transposes
        are strides of the attention kernel, not copies; the sinusoid table is added inside the LayerNorm kernel."""
        eng, w, c = self.eng, self.eng.w, self.eng.cfg
        kw = c["motion_module_kwargs"]
        heads = kw.get("num_attention_heads", 8)
        h, wd = self._geo(lvl)
        rows, hw = self._rows(lvl), h * wd
        d = cc // heads
        t = m + ".temporal_transformer"
        nrm = self.buf(rows, cc)
        self.gn([(x, cc)], rows, hw, c["norm_num_groups"], w.f32(t + ".norm.weight"), w.f32(t + ".norm.bias"), 1e-6,
                False, nrm.ptr)
        def has_pe(ab: str) -> bool:
            return w.has(ab + ".pos_encoder.pe")

        hs = self._linear(nrm, t + ".proj_in", cc,
                          stats=self._fold(lvl, cc, has_pe(f"{t}.transformer_blocks.0.attention_blocks.0")))
        del nrm
        addr = (hw, self.F * hw, 1, hw)
        i = 0
        while w.has(f"{t}.transformer_blocks.{i}.ff_norm.weight"):
            blk = f"{t}.transformer_blocks.{i}"
            k = 0
            while w.has(f"{blk}.attention_blocks.{k}.to_q.weight"):
                ab = f"{blk}.attention_blocks.{k}"
                pe = None
                if w.has(ab + ".pos_encoder.pe"):
                    name = f"{ab}#pe{self.F}"  # one table per segment length (plans with different F share the engine)
                    if name not in w.t:
                        w.put(name, w.raw(ab + ".pos_encoder.pe")[0, : self.F].contiguous())
                    pe = w.t[name]
                    assert pe.shape[0] == self.F, "more frames than temporal_position_encoding_max_len"
                n = self.buf(rows, cc)
                qkv = self.buf(rows, 3 * cc)
                if hs.aux is not None:
                    # (LN(x) + pe[frame]) W^T: the sinusoid term becomes one bias row per (batch element, frame)
                    wq, cs, bq = eng.folded(ab + ".qkv", f"{blk}.norms.{k}", lambda: torch.cat(
                        [w.raw(ab + ".to_q.weight"), w.raw(ab + ".to_k.weight"), w.raw(ab + ".to_v.weight")]),
                        pe=pe, repeat=self.B)
                    self.gemm([(hs.ptr, cc, cc, 1)], 1, 1, rows, wq, 3 * cc, qkv.ptr, 3 * cc, bias_ptr=bq.data_ptr(),
                              bias_div=hw if pe is not None else 0, bias_ld=3 * cc if pe is not None else 0,
                              col_sum_ptr=cs.data_ptr(), parts_in=self._parts(hs))
                else:
                    self.layernorm(hs.ptr, rows, cc, w.f32(f"{blk}.norms.{k}.weight"), w.f32(f"{blk}.norms.{k}.bias"),
                                   n.ptr, pe=pe, rows_per_frame=hw, nframes=self.F)
                    self.gemm([(n.ptr, cc, cc, 1)], 1, 1, rows, eng.qkv(ab), 3 * cc, qkv.ptr, 3 * cc)
                self.attention(qkv.ptr, qkv.ptr + 2 * cc, qkv.ptr + 4 * cc, n.ptr, 3 * cc, 3 * cc, 3 * cc, cc,
                               self.B * hw, heads, d, self.F, self.F, q_addr=addr, kv_addr=addr)
                del qkv
                nxt = f"{blk}.attention_blocks.{k + 1}"
                st = self._fold(lvl, cc, has_pe(nxt)) if w.has(nxt + ".to_q.weight") else self._fold(lvl, cc)
                hs2 = self._linear(n, ab + ".to_out.0", cc, residual=hs, stats=st)
                del n, hs
                hs = hs2
                k += 1
            nxt = f"{t}.transformer_blocks.{i + 1}"
            hs = self._ff(blk + ".ff_norm", blk + ".ff", hs, stats_next=w.has(nxt + ".ff_norm.weight") and self._fold(
                lvl, cc, has_pe(nxt + ".attention_blocks.0")))
            i += 1
        return self._linear(hs, t + ".proj_out", cc, residual=x, gn=self._gn_unit(lvl, cc))

    def _conv3x3(self, key: str, x: Buf, cin: int, cout: int, lvl: int) -> Buf:
        h, wd = self._geo(lvl)
        w = self.eng.w
        out = self.buf(self._rows(lvl), cout)
        gu = self._gn_unit(lvl, cout)
        self.gemm([(x.ptr, cin, cin, 9)], self.B * self.F, h, wd, w.conv(key + ".weight"), cout, out.ptr, cout,
                  bias_ptr=w.f32(key + ".bias").data_ptr(), gn_out=self.with_gn_parts(out, gu) if gu else None)
        return out

    # -- whole forward ----------------------------------------------------------------------------------------
    def _build(self) -> None:
        eng, w, c = self.eng, self.eng.w, self.eng.cfg
        B, F = self.B, self.F
        boc = c["block_out_channels"]
        nlev = len(boc)
        rows0 = self._rows(0)
        cin_pad = (c["in_channels"] + KPAD - 1) // KPAD * KPAD
        # static I/O
        self.x_in = self.static(rows0, cin_pad)
        self.t_in = self.static(1, max(B, 4), torch.float32)
        self.audio_in = self.static(B * F * self.S, eng.cross_k) if c["add_audio_layer"] else None
        self.eps_out = self.static(rows0, c["out_channels"], torch.float32)
        self.x_in.tensor().zero_()
        if self.audio_in is not None:
            self.audio_in.tensor().zero_()
        n_gn = 2 * len(eng._resnet_names()) + len(eng._attn_names()) + 64
        self.begin_stats(n_gn * B * F * 32 * 2)

        # time embedding: Timesteps -> Linear -> SiLU -> Linear (unet.py:376-382), then every resnet's
        # time_emb_proj(SiLU(emb)) + conv1.bias in one launch (resnet.py:190-205)
        te0 = self.static(B, boc[0], torch.float32)
        te1 = self.static(B, boc[0] * 4, torch.float32)
        emb = self.static(B, boc[0] * 4, torch.float32)
        self.tproj = self.static(B, eng.tproj_total, torch.float32)
        assert c["flip_sin_to_cos"] and c["freq_shift"] == 0
        n_te = len(self.ops)
        self.call("ls_timestep_embedding", self.t_in.ptr, B, boc[0], te0.ptr)
        self.call("ls_small_linear", te0.ptr, B, boc[0], w.t["te1.w"].data_ptr(),
                  w.f32("time_embedding.linear_1.bias").data_ptr(), None, boc[0] * 4, 0, 1, te1.ptr)
        self.call("ls_small_linear", te1.ptr, B, boc[0] * 4, w.t["te2.w"].data_ptr(),
                  w.f32("time_embedding.linear_2.bias").data_ptr(), None, boc[0] * 4, 0, 0, emb.ptr)
        self.call("ls_small_linear", emb.ptr, B, boc[0] * 4, w.t["tproj.w"].data_ptr(), w.t["tproj.b"].data_ptr(),
                  w.t["tproj.add"].data_ptr(), eng.tproj_total, 1, 0, self.tproj.ptr)
        self.te_ops = (n_te, len(self.ops))

        # audio K/V for all 16 cross-attention layers: one GEMM (depends only on the audio => the pipeline runs
        # this sub-plan once per segment, see UNetPlan.kv_ops)
        self.audio_kv = None
        n_before = len(self.ops)
        if self.audio_in is not None:
            self.audio_kv = self.static(B * F * self.S, eng.kv_total)
            self.gemm([(self.audio_in.ptr, eng.cross_k, eng.cross_k, 1)], 1, 1, B * F * self.S, w.t["kv.w"],
                      eng.kv_total, self.audio_kv.ptr, eng.kv_total)
        self.kv_ops = (n_before, len(self.ops))
        # neither depends on the latents: the denoising loop computes the time path for ALL its timesteps in one batched
        # pass (UNetEngine.time_table) and the audio K/V once per segment; UNet3DConditionModel.forward runs both per call
        self.hoisted = set(range(*self.te_ops)) | set(range(*self.kv_ops))

        # conv_in (unet.py:395)
        x = self.buf(rows0, boc[0])
        gu = self._gn_unit(0, boc[0])
        self.gemm([(self.x_in.ptr, cin_pad, cin_pad, 9)], B * F, self.H, self.W,
                  w.conv("conv_in.weight"), boc[0], x.ptr, boc[0], bias_ptr=w.f32("conv_in.bias").data_ptr(),
                  gn_out=self.with_gn_parts(x, gu) if gu else None)
        skips: List[Tuple[Buf, int]] = [(x, boc[0])]
        self._tap("conv_in", x, 0)
        ch = boc[0]
        for i, typ in enumerate(c["down_block_types"]):
            p = f"down_blocks.{i}"
            cout = boc[i]
            for j in range(c["layers_per_block"]):
                shared = (self.same_sample and i == 0 and j == 0 and typ == "CrossAttnDownBlock3D" and
                          c["add_audio_layer"] and self.audio_kv is not None)
                if shared:
                    # both batch elements are the same tensor up to the first audio cross-attention: one element's worth
                    # of rows (the first half of conv_in's output) through the resnet and the self-attention part
                    self.B = 1
                    x = self._resnet(f"{p}.resnets.{j}", [(x, ch)], cout, i)  # reads the first half of its input
                    self.B = B
                else:
                    x = self._resnet(f"{p}.resnets.{j}", [(x, ch)], cout, i)
                ch = cout
                if typ == "CrossAttnDownBlock3D":
                    x = self._transformer(f"{p}.attentions.{j}", x, ch, i, shared=shared)
                if w.has(f"{p}.motion_modules.{j}.temporal_transformer.norm.weight"):
                    x = self._motion(f"{p}.motion_modules.{j}", x, ch, i)
                self._tap(f"{p}.{j}", x, i)
                skips.append((x, ch))
            if i != nlev - 1:
                # Downsample3D (resnet.py:93-101): 3x3 stride 2 pad 1 via explicit im2col
                h, wd = self._geo(i)
                key = f"{p}.downsamplers.0.conv"
                y = self.buf(self._rows(i + 1), ch)
                gu = self._gn_unit(i + 1, ch)
                gno = self.with_gn_parts(y, gu) if gu else None
                if S2_INPLACE:
                    # the stride-2 convolution reads its input in place (TMA element strides): no im2col copy
                    self.gemm([(x.ptr, ch, ch, 9)], B * F, h // 2, wd // 2, w.conv(key + ".weight"), ch, y.ptr, ch,
                              bias_ptr=w.f32(key + ".bias").data_ptr(), gn_out=gno, stride2=1, stride2_pad=1)
                else:
                    cols = self.buf(self._rows(i + 1), 9 * ch)
                    self.call("ls_im2col_s2", x.ptr, B * F, h, wd, ch, cols.ptr)
                    self.gemm([(cols.ptr, 9 * ch, 9 * ch, 1)], 1, 1, self._rows(i + 1), w.conv(key + ".weight"), ch, y.ptr,
                              ch, bias_ptr=w.f32(key + ".bias").data_ptr(), gn_out=gno)
                    del cols
                x = y
                skips.append((x, ch))
        lvl = nlev - 1
        # mid (unet_blocks.py:247-260)
        x = self._resnet("mid_block.resnets.0", [(x, ch)], ch, lvl)
        x = self._transformer("mid_block.attentions.0", x, ch, lvl)
        if w.has("mid_block.motion_modules.0.temporal_transformer.norm.weight"):
            x = self._motion("mid_block.motion_modules.0", x, ch, lvl)
        x = self._resnet("mid_block.resnets.1", [(x, ch)], ch, lvl)
        self._tap("mid", x, lvl)
        rev = list(reversed(boc))
        for i, typ in enumerate(c["up_block_types"]):
            p = f"up_blocks.{i}"
            cout = rev[i]
            lvl = nlev - 1 - i
            for j in range(c["layers_per_block"] + 1):
                skip, sch = skips.pop()
                x = self._resnet(f"{p}.resnets.{j}", [(x, ch), (skip, sch)], cout, lvl)
                del skip
                ch = cout
                if typ == "CrossAttnUpBlock3D":
                    x = self._transformer(f"{p}.attentions.{j}", x, ch, lvl)
                if w.has(f"{p}.motion_modules.{j}.temporal_transformer.norm.weight"):
                    x = self._motion(f"{p}.motion_modules.{j}", x, ch, lvl)
                self._tap(f"{p}.{j}", x, lvl)
            if i != nlev - 1:
                # Upsample3D (resnet.py:47-75): nearest x2 on (h, w), then 3x3 conv
                h, wd = self._geo(lvl)
                key = f"{p}.upsamplers.0.conv"
                if UPCONV_FOLD and self._full_B * F * h * wd >= UPCONV_MIN_ROWS and wd >= 8 and ch % KPAD == 0:
                    # the upsample is folded into the convolution: four sub-pixel phase GEMMs over the low-res tensor
                    y = self.buf(self._rows(lvl - 1), ch)
                    self.upconv(x.ptr, ch, B * F, h, wd, w.upconv(key + ".weight"), ch, y.ptr,
                                w.f32(key + ".bias").data_ptr())
                    x = y
                else:
                    up = self.buf(self._rows(lvl - 1), ch)
                    self.call("ls_upsample2x", x.ptr, B * F, h, wd, ch, up.ptr)
                    x = self._conv3x3(key, up, ch, ch, lvl - 1)
                    del up
        # conv_norm_out -> SiLU -> conv_out (unet.py:464-466)
        y = self.buf(rows0, ch)
        self.gn([(x, ch)], rows0, F * self.H * self.W, c["norm_num_groups"], w.f32("conv_norm_out.weight"),
                w.f32("conv_norm_out.bias"), c["norm_eps"], True, y.ptr)
        self.gemm([(y.ptr, ch, ch, 9)], B * F, self.H, self.W, w.conv("conv_out.weight"), c["out_channels"],
                  self.eps_out.ptr, c["out_channels"], bias_ptr=w.f32("conv_out.bias").data_ptr(), flags=L.EPI_OUT_F32)
        del x, y
        assert not skips


# ------------------------------------------------------------------------------------------------- VAE engine
class VAEDecoderEngine:
    """diffusers AutoencoderKL.decode for the sd-vae-ft-mse layout (lipsync_pipeline.py:145-149)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: dict = SD_VAE_FT_MSE_CONFIG, device="cuda"):
        self.cfg = dict(cfg)
        self.device = torch.device(device)
        self.w = _Weights(state_dict, self.device)
        self.plans: Dict[Tuple[int, int, int], "VAEPlan"] = {}

    def conv_sc(self, r: str, cin: int) -> Tuple[torch.Tensor, torch.Tensor]:
        name = r + "#c2sc"
        w = self.w
        if name + ".w" not in w.t:
            w2 = pack_conv3x3(w.raw(r + ".conv2.weight"))
            ws = pack_1x1(w.raw(r + ".conv_shortcut.weight"))
            w.put(name + ".w", torch.cat([w2, ws], dim=1))
            w.put(name + ".b", w.raw(r + ".conv2.bias") + w.raw(r + ".conv_shortcut.bias"))
        return w.t[name + ".w"], w.t[name + ".b"]

    def plan(self, nimg: int, h: int, w: int) -> "VAEPlan":
        key = (nimg, h, w)
        if key not in self.plans:
            self.plans[key] = VAEPlan(self, nimg, h, w)
        return self.plans[key]


class VAEPlan(Plan):
    def __init__(self, eng: VAEDecoderEngine, nimg: int, h: int, w: int):
        super().__init__(eng.device)
        self.eng, self.nimg, self.h, self.w = eng, nimg, h, w
        self._build()

    def _gn(self, key: str, x: Buf, cc: int, hw: int, silu: bool) -> Buf:
        w = self.eng.w
        y = self.buf(x.rows, cc)
        self.gn([(x, cc)], x.rows, hw, self.eng.cfg["norm_num_groups"], w.f32(key + ".weight"), w.f32(key + ".bias"),
                1e-6, silu, y.ptr)
        return y

    def _gn_out(self, out: Buf, hw: int):
        """gn_out argument of the GEMM that writes `out` when `out` feeds a GroupNorm over hw rows per image (the
        statistics then come from that GEMM's epilogue, ls_groupnorm_parts), else None"""
        unit = out.cols // self.eng.cfg["norm_num_groups"]
        if (not GN_PARTS or hw % 128 != 0 or out.rows % 128 != 0 or unit < 1 or out.cols % unit != 0
                or not any(out.cols % bn == 0 and bn % unit == 0 for bn in range(32, 257, 32))):
            return None
        return self.with_gn_parts(out, unit)

    def _conv(self, key: str, x: Buf, cin: int, cout: int, h: int, wd: int) -> Buf:
        w = self.eng.w
        out = self.buf(x.rows, cout)
        self.gemm([(x.ptr, cin, cin, 9)], self.nimg, h, wd, w.conv(key + ".weight"), cout, out.ptr, cout,
                  bias_ptr=w.f32(key + ".bias").data_ptr(), gn_out=self._gn_out(out, h * wd))
        return out

    def _resnet(self, r: str, x: Buf, cin: int, cout: int, h: int, wd: int) -> Buf:
        """diffusers ResnetBlock2D without time embedding: GN -> SiLU -> 3x3 -> GN -> SiLU -> 3x3 (+1x1 shortcut)"""
        w = self.eng.w
        y1 = self._gn(r + ".norm1", x, cin, h * wd, True)
        h1 = self._conv(r + ".conv1", y1, cin, cout, h, wd)
        del y1
        y2 = self._gn(r + ".norm2", h1, cout, h * wd, True)
        del h1
        out = self.buf(x.rows, cout)
        gno = self._gn_out(out, h * wd)
        if w.has(r + ".conv_shortcut.weight"):
            wp, bp = self.eng.conv_sc(r, cin)
            self.gemm([(y2.ptr, cout, cout, 9), (x.ptr, cin, cin, 1)], self.nimg, h, wd, wp, cout, out.ptr, cout,
                      bias_ptr=bp.data_ptr(), gn_out=gno)
        else:
            self.gemm([(y2.ptr, cout, cout, 9)], self.nimg, h, wd, w.conv(r + ".conv2.weight"), cout, out.ptr, cout,
                      bias_ptr=w.f32(r + ".conv2.bias").data_ptr(), residual_ptr=x.ptr, ldr=cout, gn_out=gno)
        return out

    def _lin(self, x: Buf, key: str, n: int, residual: Optional[Buf] = None, gn_hw: int = 0) -> Buf:
        w = self.eng.w
        out = self.buf(x.rows, n)
        self.gemm([(x.ptr, x.cols, x.cols, 1)], 1, 1, x.rows, w.lin(key + ".weight"), n, out.ptr, n,
                  bias_ptr=w.f32(key + ".bias").data_ptr(), residual_ptr=residual.ptr if residual is not None else 0,
                  ldr=n, gn_out=self._gn_out(out, gn_hw) if gn_hw else None)
        return out

    def _mid_attention(self, a: str, x: Buf, cc: int, h: int, wd: int) -> Buf:
        """diffusers Attention in UNetMidBlock2D: GN(32, eps 1e-6) -> q,k,v Linear(with bias) -> 1 head, d = C ->
        softmax(q k^T / sqrt(C)) v -> Linear + residual"""
        hw = h * wd
        n = self._gn(a + ".group_norm", x, cc, hw, False)
        q, k, v = self._lin(n, a + ".to_q", cc), self._lin(n, a + ".to_k", cc), self._lin(n, a + ".to_v", cc)
        del n
        sc = self.buf(x.rows, hw, torch.float32)
        self.gemm([(q.ptr, cc, cc, 1)], self.nimg, 1, hw, None, hw, sc.ptr, hw, flags=L.EPI_OUT_F32,
                  b_batch_stride=hw * cc, b_ptr=k.ptr)
        del q, k
        pr = self.buf(x.rows, hw)
        self.call("ls_softmax_rows", sc.ptr, x.rows, hw, float(cc) ** -0.5, pr.ptr)
        del sc
        vt = self.buf(self.nimg * cc, hw)
        self.call("ls_transpose", v.ptr, self.nimg, hw, cc, vt.ptr)
        del v
        o = self.buf(x.rows, cc)
        self.gemm([(pr.ptr, hw, hw, 1)], self.nimg, 1, hw, None, cc, o.ptr, cc, b_batch_stride=cc * hw, b_ptr=vt.ptr)
        del pr, vt
        return self._lin(o, a + ".to_out.0", cc, residual=x, gn_hw=hw)

    def _build(self) -> None:
        eng, w, c = self.eng, self.eng.w, self.eng.cfg
        nimg, h, wd = self.nimg, self.h, self.w
        boc = list(c["block_out_channels"])
        lat = c["latent_channels"]
        rows = nimg * h * wd
        self.z_in = self.static(rows, KPAD)  # z / scaling_factor + shift, channels-last, zero padded to 64
        self.z_in.tensor().zero_()
        nup = len(boc)
        self.begin_stats((3 * 2 * nup + 12) * nimg * 32 * 2)
        # post_quant_conv (1x1, 4 -> 4) into a zero-padded 64-channel buffer
        pq = self.static(rows, KPAD)
        pq.tensor().zero_()
        self.gemm([(self.z_in.ptr, KPAD, KPAD, 1)], 1, 1, rows, w.lin("post_quant_conv.weight"), lat, pq.ptr, KPAD,
                  bias_ptr=w.f32("post_quant_conv.bias").data_ptr())
        top = boc[-1]
        x = self._conv("decoder.conv_in", pq, KPAD, top, h, wd)
        x = self._resnet("decoder.mid_block.resnets.0", x, top, top, h, wd)
        x = self._mid_attention("decoder.mid_block.attentions.0", x, top, h, wd)
        x = self._resnet("decoder.mid_block.resnets.1", x, top, top, h, wd)
        rev = list(reversed(boc))
        ch = top
        for i in range(nup):
            cout = rev[i]
            for j in range(c["layers_per_block"] + 1):
                x = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", x, ch, cout, h, wd)
                ch = cout
            if i != nup - 1:
                key = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                if UPCONV_FOLD and nimg * h * wd >= UPCONV_MIN_ROWS and wd >= 8 and ch % KPAD == 0:
                    y = self.buf(nimg * 4 * h * wd, ch)
                    self.upconv(x.ptr, ch, nimg, h, wd, w.upconv(key + ".weight"), ch, y.ptr, w.f32(key + ".bias").data_ptr())
                    h, wd = 2 * h, 2 * wd
                    x = y
                else:
                    up = self.buf(nimg * 4 * h * wd, ch)
                    self.call("ls_upsample2x", x.ptr, nimg, h, wd, ch, up.ptr)
                    h, wd = 2 * h, 2 * wd
                    x = self._conv(key, up, ch, ch, h, wd)
                    del up
        y = self._gn("decoder.conv_norm_out", x, ch, h * wd, True)
        self.out_h, self.out_w = h, wd
        nout = c["out_channels"]
        self.ld_out = 4
        self.dec_out = self.static(nimg * h * wd, self.ld_out, torch.float32)
        self.gemm([(y.ptr, ch, ch, 9)], nimg, h, wd, w.conv("decoder.conv_out.weight"), nout, self.dec_out.ptr,
                  self.ld_out, bias_ptr=w.f32("decoder.conv_out.bias").data_ptr(), flags=L.EPI_OUT_F32)
        del x, y


# ------------------------------------------------------------------------------------------ VAE encoder engine
class VAEEncoderEngine(VAEDecoderEngine):
    """diffusers AutoencoderKL.encode(x).latent_dist.parameters for the sd-vae-ft-mse layout (call sites
    lipsync_pipeline.py:298,315).  Same kernels as the decoder: implicit-GEMM 3x3 convolutions,
    GroupNorm + SiLU passes, single-head mid attention; the stride-2 Downsample2D(padding=0) goes through an explicit
    im2col with the library's asymmetric (0, 1, 0, 1) zero padding."""

    def plan(self, nimg: int, h: int, w: int) -> "VAEEncodePlan":
        key = (nimg, h, w)
        if key not in self.plans:
            self.plans[key] = VAEEncodePlan(self, nimg, h, w)
        return self.plans[key]


class VAEEncodePlan(VAEPlan):
    """x_in: fp16 channels-last [(n H W), 64] pixels in [-1, 1] (3 real channels); mom_out: fp32 [(n h w), 8] =
    [mean | logvar] of the diagonal Gaussian (h = H/8)."""

    def _build(self) -> None:
        eng, w, c = self.eng, self.eng.w, self.eng.cfg
        nimg, h, wd = self.nimg, self.h, self.w
        boc = list(c["block_out_channels"])
        lat = c["latent_channels"]
        self.x_in = self.static(nimg * h * wd, KPAD)
        self.x_in.tensor().zero_()
        self.begin_stats((2 * c["layers_per_block"] * len(boc) + 12) * nimg * 32 * 2)
        x = self._conv("encoder.conv_in", self.x_in, KPAD, boc[0], h, wd)
        ch = boc[0]
        for i, cout in enumerate(boc):
            for j in range(c["layers_per_block"]):
                x = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", x, ch, cout, h, wd)
                ch = cout
            if i != len(boc) - 1:
                key = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                rows = nimg * (h // 2) * (wd // 2)
                y = self.buf(rows, ch)
                if S2_INPLACE:
                    h, wd = h // 2, wd // 2
                    self.gemm([(x.ptr, ch, ch, 9)], nimg, h, wd, w.conv(key + ".weight"), ch, y.ptr, ch,
                              bias_ptr=w.f32(key + ".bias").data_ptr(), gn_out=self._gn_out(y, h * wd), stride2=1,
                              stride2_pad=0)
                else:
                    cols = self.buf(rows, 9 * ch)
                    self.call("ls_im2col_s2_pad", x.ptr, nimg, h, wd, ch, 0, cols.ptr)
                    h, wd = h // 2, wd // 2
                    self.gemm([(cols.ptr, 9 * ch, 9 * ch, 1)], 1, 1, rows, w.conv(key + ".weight"), ch, y.ptr, ch,
                              bias_ptr=w.f32(key + ".bias").data_ptr(), gn_out=self._gn_out(y, h * wd))
                    del cols
                x = y
        x = self._resnet("encoder.mid_block.resnets.0", x, ch, ch, h, wd)
        x = self._mid_attention("encoder.mid_block.attentions.0", x, ch, h, wd)
        x = self._resnet("encoder.mid_block.resnets.1", x, ch, ch, h, wd)
        y = self._gn("encoder.conv_norm_out", x, ch, h * wd, True)
        rows = nimg * h * wd
        # conv_out (3x3, C -> 2*latent) into a zero-padded 64-channel buffer, then the 1x1 quant_conv in fp32
        co = self.static(rows, KPAD)
        co.tensor().zero_()
        self.gemm([(y.ptr, ch, ch, 9)], nimg, h, wd, w.conv("encoder.conv_out.weight"), 2 * lat, co.ptr, KPAD,
                  bias_ptr=w.f32("encoder.conv_out.bias").data_ptr())
        self.out_h, self.out_w = h, wd
        self.ld_out = 2 * lat
        self.mom_out = self.static(rows, self.ld_out, torch.float32)
        self.gemm([(co.ptr, KPAD, KPAD, 1)], 1, 1, rows, w.lin("quant_conv.weight"), 2 * lat, self.mom_out.ptr,
                  self.ld_out, bias_ptr=w.f32("quant_conv.bias").data_ptr(), flags=L.EPI_OUT_F32)
        del x, y

