// Sub-daily disaggregation: C ABI entry points (argument checks, kernel selection).
// The kernels are in disagg_event.cu and disagg_interior.cu; disagg_common.cuh has the map.
#include "disagg_common.cuh"

using namespace msb;

extern "C" size_t msb_disagg_workspace_bytes(int64_t ld, int day0, int day1) {
  if (ld <= 0 || day1 <= day0) return 0;
  const size_t n_rec = (size_t)(day1 - day0) + 2 * MSB_REC_MARGIN;
  return (size_t)MSB_REC_FIELDS * n_rec * (size_t)ld * sizeof(double);
}

extern "C" size_t msb_disagg_workspace_bytes_lw(int64_t ld, int day0, int day1, int time_step) {
  if (ld <= 0 || day1 <= day0 || time_step <= 0 || time_step >= kMinPerDay) return 0;
  const size_t rec = (msb_disagg_workspace_bytes(ld, day0, day1) + 255) & ~(size_t)255;
  return rec + (size_t)(day1 - day0) * (size_t)(kMinPerDay / time_step) * (size_t)ld * sizeof(double);
}

extern "C" int msb_disaggregate(const msb_params *p, const msb_forcing *f,
                                const msb_solar_tables *t, const msb_daily *daily,
                                const msb_subdaily *out, void *stream) {
  if (!p || !f || !t || !daily || !out) { set_error("msb_disaggregate: NULL argument"); return MSB_E_ARG; }
  if (f->n_cells <= 0 || f->n_days < 2 || f->ld < f->n_cells || out->ld < f->n_cells) {
    set_error("msb_disaggregate: bad shapes");
    return MSB_E_ARG;
  }
  const int ts = p->time_step;
  if (ts <= 0 || ts >= kMinPerDay || kMinPerDay % ts || ts > 360) {
    set_error("Time step must be evenly divisible into 1440 minutes (24 hours) and less than 360 "
              "minutes (6 hours). Got %d.", ts);   // metsim.py:641-646
    return MSB_E_ARG;
  }
  if (p->prec_type != MSB_PREC_UNIFORM && out->var[MSB_V_PREC] && (!f->t_pk12 || !f->dur12)) {
    set_error("msb_disaggregate: triangle/mix precipitation needs t_pk12 and dur12");
    return MSB_E_ARG;
  }
  if (out->dtype != 4 && out->dtype != 8) { set_error("msb_disaggregate: out dtype must be 4 or 8"); return MSB_E_ARG; }
  if (out->day0 < 0 || out->day1 > f->n_days || out->day0 >= out->day1) {
    set_error("msb_disaggregate: bad day range [%d, %d)", out->day0, out->day1);
    return MSB_E_ARG;
  }
  const double *vp = daily->var[MSB_D_VAPOR_PRESSURE], *sw = daily->var[MSB_D_SHORTWAVE],
               *tk = daily->var[MSB_D_TSKC];
  const bool single = msb_daily_single_pass(p, f, daily) != 0;   // candidate planes instead (daily.cu)
  if (single) {
    vp = daily->cand;
    sw = daily->cand + (size_t)f->n_days * (size_t)f->ld;
  }
  if (!vp || !sw || !tk || !daily->status) {
    set_error("msb_disaggregate: daily vapor_pressure / shortwave / tskc / status are required");
    return MSB_E_ARG;
  }
  DisaggArgs a;
  a.p = *p; a.f = *f; a.t = *t; a.vp_d = vp; a.sw_d = sw; a.tskc_d = tk; a.status = daily->status;
  a.cand_sel = single ? daily->cand_sel : nullptr;
  a.cand_stride = single ? 2 * (size_t)f->n_days * (size_t)f->ld : 0;
  a.out = *out;
  a.tpd = kMinPerDay / ts;
  const int n_days = out->day1 - out->day0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int cell_blocks = (f->n_cells + kDisaggThreads - 1) / kDisaggThreads;
  // day records of the padded days around the range (-1 and D are the padded PCHIP ends)
  a.e0 = out->day0 - MSB_REC_MARGIN < -1 ? -1 : out->day0 - MSB_REC_MARGIN;
  const int e1 = out->day1 + MSB_REC_MARGIN > f->n_days + 1 ? f->n_days + 1 : out->day1 + MSB_REC_MARGIN;
  a.n_rec = e1 - a.e0;
  a.rec = reinterpret_cast<double *>(out->workspace);
  if (!out->workspace ||
      out->workspace_bytes < (size_t)MSB_REC_FIELDS * a.n_rec * (size_t)f->ld * sizeof(double)) {
    set_error("msb_disaggregate: workspace of msb_disagg_workspace_bytes() bytes is required");
    return MSB_E_ARG;
  }
  a.fill[0] = a.fill[1] = a.fill[2] = a.fill[3] = 0;   // no valid records yet (launch_day_records)
  a.sw_full_day = p->method == MSB_METHOD_PASSTHROUGH && !p->sw_daylight;
  a.lw_raw = nullptr;
  if (p->method == MSB_METHOD_PASSTHROUGH && f->given_longwave && out->var[MSB_V_LONGWAVE]) {
    const size_t rec_bytes = (msb_disagg_workspace_bytes(f->ld, out->day0, out->day1) + 255) & ~(size_t)255;
    if (out->workspace_bytes < rec_bytes + (size_t)n_days * a.tpd * (size_t)out->ld * sizeof(double)) {
      set_error("msb_disaggregate: the longwave rescale needs a workspace of "
                "msb_disagg_workspace_bytes_lw() bytes");
      return MSB_E_ARG;
    }
    a.lw_raw = reinterpret_cast<double *>(reinterpret_cast<char *>(out->workspace) + rec_bytes);
  }
  const bool fast_lw = p->lw_type == MSB_LW_PRATA && p->lw_cloud == MSB_CLOUD_DEARDORFF;
  bool units = false;
  a.var_mask = 0;
  for (int v = 0; v < MSB_N_SUBVARS; ++v) {
    if (!out->var[v] || (v == MSB_V_WIND && !f->wind)) continue;
    a.var_mask |= 1u << v;
    if (out->scale[v] != 1.0 || out->offset[v] != 0.0) units = true;
  }
  if (!a.var_mask) { set_error("msb_disaggregate: no output variable requested"); return MSB_E_ARG; }
  a.units = units;
  a.fast_lw = fast_lw;
  // Every variable set, with or without unit conversion, with any longwave option, takes the
  // specialised kernels on the interior days; `full` is the default set
  // (examples/example_nc.conf) at float32 without conversion, which has unconditional-store
  // instantiations and, for the edge days, the two compile-time halves of disagg_kernel.
  // msb_tuning.disagg_path = 1 forces the event-driven kernel for every day (the parity tests use it
  // to keep both paths covered).
  const bool example = out->tune.disagg_path != 1;
  const bool f4 = out->dtype == 4;
  const bool full = example && a.var_mask == kExampleMask && f4 && !units && fast_lw && !a.lw_raw;
  const unsigned knot_bits = kThermoBits | (1u << MSB_V_WIND) | (1u << MSB_V_TSKC);   // thermo_knot_kernel's outputs
  const unsigned day_bits = (1u << MSB_V_PREC) | (1u << MSB_V_SHORTWAVE);  // sw_day_kernel, prec_day_kernel
  a.inv_sw_den = 1.0 / (3600.0 * ((double)ts / 60.0));
  const int D = f->n_days;
  // interior days [a0, b0) of this call: every knot exists, no NaN ends, no wrapped stream
  const int a0 = out->day0 > 2 ? out->day0 : 2, b0 = out->day1 < D - 2 ? out->day1 : D - 2;
  int rc = MSB_OK;
#define MSB_TRY(expr) do { rc = (expr); if (rc != MSB_OK) return rc; } while (0)
  if (example && a0 < b0) {
    // the record's first / last two days: day records + disagg_kernel on that sub-range (output
    // pointers shifted); everything else: thermo_knot_kernel + sw_day_kernel + prec_day_kernel, no records
    // (all records first: with the utc roll the first day of a record reads the last one's)
    if (out->day0 < a0) MSB_TRY(launch_day_records(a, out->day0 - MSB_REC_MARGIN, a0 + MSB_REC_MARGIN, e1, cell_blocks, st));
    if (b0 < out->day1) MSB_TRY(launch_day_records(a, b0 - MSB_REC_MARGIN, out->day1 + MSB_REC_MARGIN, e1, cell_blocks, st));
    auto launch_edge = [&](int s0, int s1) -> int {
      if (s0 >= s1) return MSB_OK;
      int r;
      DisaggArgs b = a;
      b.out.day0 = s0; b.out.day1 = s1;
      for (int v = 0; v < MSB_N_SUBVARS; ++v)
        if (b.out.var[v])
          b.out.var[v] = reinterpret_cast<char *>(b.out.var[v]) +
                         (size_t)(s0 - out->day0) * a.tpd * (size_t)out->ld * (size_t)out->dtype;
      if (b.lw_raw) b.lw_raw += (size_t)(s0 - out->day0) * a.tpd * (size_t)out->ld;
      if (!full) return launch_event_kernel(b, f4 ? kEvAnyF4 : kEvAnyF8, fast_lw, cell_blocks, st);
      r = launch_event_kernel(b, kEvThermoF4, true, cell_blocks, st);
      return r != MSB_OK ? r : launch_event_kernel(b, kEvStreamsF4, true, cell_blocks, st);
    };
    MSB_TRY(launch_edge(out->day0, a0));
    MSB_TRY(launch_edge(b0, out->day1));
    if (a.var_mask & knot_bits) MSB_TRY(launch_interior_thermo(a, a0, b0, cell_blocks, st));
    if (a.var_mask & day_bits) MSB_TRY(launch_interior_streams(a, a0, b0, cell_blocks, st));
  } else {
    MSB_TRY(launch_day_records(a, a.e0, e1, e1, cell_blocks, st));
    if (full) {            // too short for interior days: the thermodynamic half and the stream half
      MSB_TRY(launch_event_kernel(a, kEvThermoF4, true, cell_blocks, st));
      MSB_TRY(launch_event_kernel(a, kEvStreamsF4, true, cell_blocks, st));
    } else {
      MSB_TRY(launch_event_kernel(a, f4 ? kEvAnyF4 : kEvAnyF8, fast_lw, cell_blocks, st));
    }
  }
  if (a.lw_raw) MSB_TRY(launch_lw_rescale(a, cell_blocks, st));
#undef MSB_TRY
  return MSB_OK;
}
