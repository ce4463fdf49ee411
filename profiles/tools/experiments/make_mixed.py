"""Generates profiles/tools/experiments/mixed_interior.cu from the PRODUCT sources: one kernel whose CTAs
take one of three roles -- the thermodynamic knot kernel, the shortwave day kernel, the precipitation
day kernel, bodies unchanged -- interleaved in launch order (T, S, P, T, S, P ...) so that every SM
holds warps of different instruction mixes at the same time (the block scheduler does not co-schedule
two grids: profiles/README.md).  Build with

    python profiles/tools/experiments/make_mixed.py
    X="SRC:disagg_interior.cu=profiles/tools/experiments/mixed_interior.cu SRC:disagg_streams.cu=profiles/tools/experiments/empty.cu"
    python profiles/tools/build_variants.py mixed6="$X -DMSB_MIXED_MINB=6" ...
"""
import os
import re
import sys

ROLES = sys.argv[1] if len(sys.argv) > 1 else "TSP"     # which kernels share the mixed launch: TSP, TS or TP
OUTNAME = "mixed_interior.cu" if ROLES == "TSP" else f"mixed_interior_{ROLES.lower()}.cu"

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "..", "metsim_b200", "csrc")
interior = open(os.path.join(CSRC, "disagg_interior.cu")).read()
streams = open(os.path.join(CSRC, "disagg_streams.cu")).read()

# ---- streams: body takes the block indices; product launch function renamed ----
old = "__device__ __forceinline__ void streams_body(const DisaggArgs &a) {"
assert streams.count(old) == 1
a = streams.index(old)
b = streams.index("// latency-bound: 64 registers", a)
body = streams[a:b].replace(old, "__device__ __forceinline__ void streams_body(const DisaggArgs &a, const int bx, const int by) {")
assert body.count("blockIdx.y") == 1 and body.count("blockIdx.x") == 1
body = body.replace("blockIdx.y", "by").replace("blockIdx.x", "bx")
streams = streams[:a] + body + streams[b:]
n = streams.count("(a); }")
streams = re.sub(r"(streams_body<[^>]*>)\(a\); }", r"\1(a, blockIdx.x, blockIdx.y); }", streams)
assert "streams_body<OutT, UNITS, kStreamPrec>(a, blockIdx.x, blockIdx.y)" in streams
streams = streams.replace("int launch_interior_streams(", "int launch_interior_streams_product(")
streams = streams.replace('#include "disagg_common.cuh"', "// (disagg_common.cuh comes with the interior part)")
streams = streams.replace("}  // namespace msb", "}  // namespace msb (streams part)")

# ---- interior: kernel body -> device function ----
old = ("template <typename OutT, int MINB, bool MASKED>\n__global__ void __launch_bounds__(kDisaggThreads, MINB)\n"
       "thermo_knot_kernel(const __grid_constant__ DisaggArgs a) {")
assert interior.count(old) == 1
a = interior.index(old)
b = interior.index("// accuracy probe of the hand-written fp64 sequences", a)
body = interior[a:b].replace(old, "template <typename OutT, bool MASKED>\n__device__ __forceinline__ void "
                                  "thermo_body(const DisaggArgs &a, const int bx, const int by) {")
assert body.count("blockIdx.y") == 1 and body.count("blockIdx.x") == 1
body = body.replace("blockIdx.y", "by").replace("blockIdx.x", "bx")
wrapper = '''template <typename OutT, int MINB, bool MASKED>
__global__ void __launch_bounds__(kDisaggThreads, MINB)
thermo_knot_kernel(const __grid_constant__ DisaggArgs a) { thermo_body<OutT, MASKED>(a, blockIdx.x, blockIdx.y); }

}  // namespace msb
// ======================= product disagg_streams.cu, body parameterised =======================
''' + streams + '''
namespace msb {
// ======================= the mixed kernel =======================
#ifndef MSB_MIXED_MINB
#define MSB_MIXED_MINB 6
#endif
template <typename OutT>
__global__ void __launch_bounds__(kDisaggThreads, MSB_MIXED_MINB)
interior_mixed_kernel(const __grid_constant__ DisaggArgs aT, const __grid_constant__ DisaggArgs aS,
                      const int cell_blocks, const int n_ty, const int n_sy) {
  const unsigned b = blockIdx.x;
  const int role = b % @N@u;
  const unsigned q = b / @N@u;
  const int bx = q % (unsigned)cell_blocks, by = q / (unsigned)cell_blocks;
  if (role == 0) {
    if (by < n_ty) thermo_body<OutT, false>(aT, bx, by);
  } else if (role == 1) {
    if (by < n_sy) streams_body<OutT, false, SECOND_ROLE>(aS, bx, by);
  } else {
    if (by < n_sy) streams_body<OutT, false, kStreamPrec>(aS, bx, by);
  }
}
static thread_local unsigned g_streams_done = 0;     // output bits the mixed kernel of this call has taken

'''
interior = interior[:a] + body + wrapper + interior[b:]

# ---- launch: the mixed kernel when the request is the benchmarked one ----
old = "  const dim3 grid(cell_blocks, (a.k.e1 - a.k.e0 + a.k.tile - 1) / a.k.tile);\n  const bool masked ="
assert interior.count(old) == 1
new = '''  const dim3 grid(cell_blocks, (a.k.e1 - a.k.e0 + a.k.tile - 1) / a.k.tile);
  {
    const bool want_sw = a.var_mask & (1u << MSB_V_SHORTWAVE), want_pr = a.var_mask & (1u << MSB_V_PREC);
    const bool win = a.t.q_win && a.t.q_win_step == a.p.time_step;
    const bool plain = a.var_mask == kExampleMask && !a.units && a.fast_lw && !a.lw_raw && want_sw && want_pr && !win;
    if (plain) {
      DisaggArgs s = a;                       // the streams' work range, as launch_interior_streams sets it
      s.k.i_lo = s.k.i_hi = 0;
      s.k.e0 = a0; s.k.e1 = b0;
      s.k.tile = kStreamTile;
      while (s.k.tile > 2 && (long)cell_blocks * ((b0 - a0 + s.k.tile - 1) / s.k.tile) < 148L * 8 * 6) s.k.tile >>= 1;
      if (a.out.tune.day_tile > 0) s.k.tile = a.out.tune.day_tile < kDayWin - 8 ? a.out.tune.day_tile : kDayWin - 8;
      s.k.pf = a.out.tune.day_prefetch < 0 ? 0 : (a.out.tune.day_prefetch == 1 ? 1 : 2);
      const int n_ty = grid.y, n_sy = (b0 - a0 + s.k.tile - 1) / s.k.tile;
      const unsigned total = @N@u * (unsigned)cell_blocks * (unsigned)(n_ty > n_sy ? n_ty : n_sy);
      if (a.out.dtype == 4) interior_mixed_kernel<float><<<total, kDisaggThreads, 0, st>>>(a, s, cell_blocks, n_ty, n_sy);
      else interior_mixed_kernel<double><<<total, kDisaggThreads, 0, st>>>(a, s, cell_blocks, n_ty, n_sy);
      MSB_CUDA(cudaGetLastError());
      g_streams_done = @BITS@;
      return MSB_OK;
    }
  }
  const bool masked ='''
interior = interior.replace(old, new)
old = "}  // namespace msb\n\nusing namespace msb;"
assert interior.count(old) == 1
interior = interior.replace(old, '''int launch_interior_streams(DisaggArgs &a, int a0, int b0, int cell_blocks, cudaStream_t st) {
  const unsigned taken = g_streams_done, keep = a.var_mask;
  g_streams_done = 0;
  a.var_mask &= ~taken;                                              // what the mixed kernel of this call took
  const int rc = (a.var_mask & ((1u << MSB_V_PREC) | (1u << MSB_V_SHORTWAVE)))
                     ? launch_interior_streams_product(a, a0, b0, cell_blocks, st) : MSB_OK;
  a.var_mask = keep;
  return rc;
}

}  // namespace msb

using namespace msb;''')
n_roles = len(ROLES)
second = "kStreamSw" if ROLES[1] == "S" else "kStreamPrec"
bits = {"TSP": "(1u << MSB_V_PREC) | (1u << MSB_V_SHORTWAVE)", "TS": "(1u << MSB_V_SHORTWAVE)", "TP": "(1u << MSB_V_PREC)"}[ROLES]
interior = interior.replace("@N@", str(n_roles)).replace("SECOND_ROLE", second).replace("@BITS@", bits)
open(os.path.join(HERE, OUTNAME), "w").write(
    "// GENERATED by make_mixed.py from metsim_b200/csrc/disagg_interior.cu + disagg_streams.cu -- experiment only\n" + interior)
open(os.path.join(HERE, "empty.cu"), "w").write("// stands in for disagg_streams.cu in the mixed-kernel experiment (its code is part of mixed_interior.cu)\n")
print("written")
