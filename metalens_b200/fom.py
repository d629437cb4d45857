"""Far-field figure of merit and CUDA-graph replay for parameter sweeps.

The reference's ``vary_angle`` / ``optimize`` loops (grating.py:685-918) score a candidate with a
figure of merit computed inside the S4 solver (grating.lua:188-253); NF->FF is never called there.
BASELINE config 5 asks for the new composition "per-step NF->FF figure of merit inside the loop":
one small aperture per step, transformed and reduced to a scalar.  Such steps are launch-latency
bound (a 256x256 aperture is microseconds of work), so the whole step -- aperture sums, power
epilogue, cone reduction -- is captured once in a CUDA graph and replayed per step; the caller only
rewrites the static input buffers.

    FOM = sum of P over the cone |u - u_target| <= half_width  /  sum of P over all finite bins
"""
import ctypes as C

import torch

from . import _lib


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FarfieldFOM:
    """Couples a FarfieldPlan with the cone-power reduction; optionally graph-captured."""

    def __init__(self, plan, target_ux, target_uy, half_width):
        self.plan = plan
        self.lib = plan.lib
        self.target = [float(target_ux), float(target_uy), float(half_width)]
        dev = plan.device
        self.nb = self.lib.mlb_ff_epilogue_blocks(plan.Kx, plan.Ky)
        self.cone_sums = torch.empty(self.nb, dtype=torch.float64, device=dev)
        self.total_sums = torch.empty(self.nb, dtype=torch.float64, device=dev)
        self.result = torch.zeros(2, dtype=torch.float64, device=dev)        # [cone power, total power]
        # static inputs for graph replay: callers write new apertures into these
        self.static_fields = torch.zeros((4, plan.Mx, plan.My + (plan.My & 1)), dtype=torch.complex64, device=dev)
        self.graph = None

    def set_target(self, target_ux, target_uy, half_width=None):
        """Move the cone (invalidates a captured graph: kernel arguments are baked in)."""
        self.target = [float(target_ux), float(target_uy), self.target[2] if half_width is None else float(half_width)]
        self.graph = None

    def _fields(self):
        return [self.static_fields[i][:, :self.plan.My] for i in range(4)]

    def _launch(self):
        p = self.plan
        p.run(self._fields())
        rc = self.lib.mlb_cone_power(p.P.data_ptr(), p.P.shape[1], 1 if p.P.dtype == torch.float64 else 0,
                                     p.d_ux.data_ptr(), p.d_uy.data_ptr(), p.Kx, p.Ky, self.target[0], self.target[1],
                                     self.target[2], self.cone_sums.data_ptr(), self.total_sums.data_ptr(),
                                     _stream_ptr())
        _lib.check(rc, "mlb_cone_power")
        scale = p.dux * p.duy
        _lib.check(self.lib.mlb_sum_f64(self.cone_sums.data_ptr(), self.nb, scale, self.result.data_ptr(),
                                        _stream_ptr()), "mlb_sum_f64")
        _lib.check(self.lib.mlb_sum_f64(self.total_sums.data_ptr(), self.nb, scale,
                                        self.result.data_ptr() + 8, _stream_ptr()), "mlb_sum_f64")

    def evaluate(self, fields=None):
        """Eager evaluation.  `fields`: 4 device complex64 (Mx,My) tensors copied into the static
        buffers, or None to use what is already there.  Returns the device tensor [cone, total]."""
        if fields is not None:
            for i, f in enumerate(fields):
                self.static_fields[i][:, :self.plan.My].copy_(f)
        self._launch()
        return self.result

    def capture(self):
        """Capture one step (all kernels of run() + the reductions) into a CUDA graph."""
        self._launch()                                   # warm-up outside capture (lazy attribute sets)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch()
        self.graph = g
        return g

    def replay(self):
        """Replay the captured step on whatever is in `static_fields`; returns [cone, total]."""
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.result

    @staticmethod
    def fom(result):
        r = result.detach().cpu()
        return float(r[0] / r[1]) if float(r[1]) != 0 else float("nan")
