// Host-side file formats either side of the hot path: the reference's params.in, SAC traces,
// reference velocity model, and the mcmc_out output set.  Plain C++; no CUDA calls in this file.
//
//   rfinv_problem_load     <- get_params (src/params.f90:101-388), get_line (:392-405), read_obs (:422-476),
//                             read_ref_model (src/model.f90:109-171)
//   rfinv_write_outputs    <- output_results (src/mcmc_out.f90:97-318)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "rfinv_common.cuh"

struct rfinv_problem {
  rfinv_config cfg;
  std::string out_dir, vel_file, base_dir;
  std::vector<std::string> obs_files;
  std::vector<double> rayps, a_gus, obs, vp_ref, vs_ref, sig_min, sig_max;
  std::vector<int32_t> ipha;
  double t_end = 0.0;
  std::vector<std::string> raw;  // the non-comment lines, in order (params.in.copy)
};

namespace {

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

// get_line (src/params.f90:392-405): skip lines whose first non-blank character is '#'.  (The reference reads into
// a character(100) buffer: longer lines are truncated there; we keep them whole.)
bool next_line(std::istream& in, std::string& out) {
  std::string line;
  while (std::getline(in, line)) {
    std::string t = trim(line);
    if (!t.empty() && t[0] == '#') continue;
    out = t;
    return true;
  }
  return false;
}

// list-directed read of a character variable: quoted ('...' or "...") or a blank-delimited token
std::string parse_string(const std::string& s) {
  if (s.empty()) return s;
  if (s[0] == '\'' || s[0] == '"') {
    size_t e = s.find(s[0], 1);
    return s.substr(1, e == std::string::npos ? std::string::npos : e - 1);
  }
  std::istringstream is(s);
  std::string tok;
  is >> tok;
  return tok;
}

// list-directed numeric read: blanks or commas separate values; Fortran 'd' exponents accepted
bool parse_numbers(const std::string& s, int n, double* out) {
  std::string t = s;
  for (char& ch : t) {
    if (ch == ',') ch = ' ';
    if (ch == 'd' || ch == 'D') ch = 'e';
  }
  std::istringstream is(t);
  for (int i = 0; i < n; ++i)
    if (!(is >> out[i])) return false;
  return true;
}

std::string join_path(const std::string& base, const std::string& p) {
  if (p.empty() || p[0] == '/' || base.empty()) return p;
  return base + "/" + p;
}

int fail(int code, const char* fmt, const std::string& a) {
  rfinv_set_error(fmt, a.c_str());
  return code;
}

// read_obs (src/params.f90:422-476): SAC binary, native endianness, 4-byte records: 1 = delta, 6 = b, 80 = npts,
// data from record 159.  Window: it1 = nint((t_start-b)/delta)+1, nsmp = it2-it1+1.
int read_sac_window(const std::string& path, double t_start, double t_end, int& nsmp, double& delta, std::vector<double>& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(RFINV_ERR_IO, "ERROR: cannot open %s", path);
  f.seekg(0, std::ios::end);
  const long long nbytes = f.tellg();
  std::vector<float> raw((size_t)(nbytes / 4));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)(raw.size() * 4));
  if (raw.size() < 158) return fail(RFINV_ERR_IO, "ERROR: %s is too short for a SAC header", path);
  const float delta4 = raw[0], t_beg4 = raw[5];
  const int it1 = f_nint((t_start - (double)t_beg4) / (double)delta4) + 1;
  const int it2 = f_nint((t_end - (double)t_beg4) / (double)delta4) + 1;
  nsmp = it2 - it1 + 1;
  delta = (double)delta4;
  if (nsmp < 1 || it1 < 1 || (size_t)(158 + it2) > raw.size())
    return fail(RFINV_ERR_IO, "ERROR: time window [T_START, T_END] is outside %s", path);
  out.resize(nsmp);
  for (int it = 1; it <= nsmp; ++it) out[it - 1] = (double)raw[158 + it + it1 - 1 - 1];  // record 158+it+it1-1, 1-based
  return RFINV_OK;
}

// read_ref_model (src/model.f90:109-171): text "depth vp vs", constant depth increment (checked to 1.0e-5)
int read_velmod(const std::string& path, rfinv_problem* p) {
  std::ifstream f(path);
  if (!f) return fail(RFINV_ERR_IO, "ERROR: cannot open %s", path);
  std::string line;
  std::vector<double> z;
  while (std::getline(f, line)) {
    double v[3];
    if (!parse_numbers(trim(line), 3, v)) break;
    z.push_back(v[0]); p->vp_ref.push_back(v[1]); p->vs_ref.push_back(v[2]);
  }
  if (z.empty()) return fail(RFINV_ERR_IO, "ERROR: no layers in %s", path);
  p->cfg.z_ref_min = z[0];
  double dz = -100.0, z_old = -999.0;
  for (size_t i = 0; i < z.size(); ++i) {
    if (i + 1 >= 3 && std::fabs(z[i] - z_old - dz) > (double)1.0e-5f)
      return fail(RFINV_ERR_IO, "ERROR: Depth increment must be constant in %s", path);
    dz = z[i] - z_old;
    z_old = z[i];
  }
  p->cfg.dz_ref = dz;
  p->cfg.nref = (int)z.size();
  return RFINV_OK;
}

// gfortran list-directed output of a real(8): a 26-column field; inside [0.1, 1e17) an F edit descriptor with 17
// significant digits right-aligned in the first 21 columns (then 5 blanks), otherwise d.dddddddddddddddddE+ddd.
std::string ld_real(double x) {
  char buf[80];
  const double ax = std::fabs(x);
  std::string body;
  bool fixed = true;
  if (x == 0.0) {
    body = "0.0000000000000000";
  } else if (ax >= 0.1 && ax < 1e17) {
    const int e = (int)std::floor(std::log10(ax)) + 1;       // digits before the decimal point (0 for 0.1 <= |x| < 1)
    const int dec = e >= 1 ? 17 - e : 17;
    snprintf(buf, sizeof buf, "%.*f", dec < 0 ? 0 : dec, x);
    body = buf;
  } else if (std::isfinite(x)) {
    fixed = false;
    snprintf(buf, sizeof buf, "%.16E", x);
    std::string t(buf);
    const size_t pos = t.find('E');
    char eb[16];
    snprintf(eb, sizeof eb, "E%c%03d", t[pos + 1], std::abs(std::atoi(t.c_str() + pos + 2)));
    body = t.substr(0, pos) + eb;
  } else {
    body = std::isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity");
  }
  const int width = fixed ? 21 : 26;
  std::string out = (int)body.size() < width ? std::string(width - body.size(), ' ') + body : " " + body;
  if (fixed) out += "     ";
  return out;
}
std::string ld_int(long long v) {
  char buf[32];
  snprintf(buf, sizeof buf, "%12lld", v);
  return buf;
}

}  // namespace

extern "C" {

// Parses params.in (48 values in fixed order), reads the SAC traces (window [T_START, T_END]) and the reference velocity
// model.  Relative paths are resolved against `base_dir` (NULL/"" = current directory, like the reference).
int32_t rfinv_problem_load(const char* params_path, const char* base_dir, rfinv_problem** out) {
  if (!params_path || !out) { rfinv_set_error("rfinv_problem_load: NULL argument"); return RFINV_ERR_ARG; }
  *out = nullptr;
  std::ifstream in(params_path);
  if (!in) return fail(RFINV_ERR_IO, "ERROR: cannot open : %s", params_path);
  rfinv_problem* p = new rfinv_problem();
  std::memset(&p->cfg, 0, sizeof(p->cfg));
  p->base_dir = base_dir ? base_dir : "";
  rfinv_config& c = p->cfg;
  std::string line;
  double v[4];
  int st = RFINV_OK;
#define NEXT(what)                                                                                   \
  if (!next_line(in, line)) { st = fail(RFINV_ERR_IO, "ERROR: params.in ended while reading %s", what); goto done; } \
  p->raw.push_back(line);
#define NUM(n, what)                                                                                 \
  NEXT(what)                                                                                         \
  if (!parse_numbers(line, n, v)) { st = fail(RFINV_ERR_IO, "ERROR: cannot parse %s", what); goto done; }
  NEXT("OUT_DIR") p->out_dir = parse_string(line);
  NUM(1, "N_BURN") c.nburn = (int)v[0];
  NUM(1, "N_ITER") c.niter = (int)v[0];
  NUM(1, "N_CORR") c.ncorr = (int)v[0];
  NUM(1, "N_CHAINS") c.nchains = (int)v[0];
  NUM(1, "N_COOL") c.ncool = (int)v[0];
  NUM(1, "T_HIGH") c.t_high = v[0];
  NUM(1, "I_SEED") c.iseed = (int)v[0];
  NUM(1, "N_TRC") c.ntrc = (int)v[0];
  if (c.ntrc < 1 || c.ntrc > RFINV_MAX_TRC) { st = fail(RFINV_ERR_ARG, "ERROR: N_TRC out of range%s", ""); goto done; }
  for (int t = 0; t < c.ntrc; ++t) { NUM(1, "RAYP") p->rayps.push_back(v[0]); }
  for (int t = 0; t < c.ntrc; ++t) { NUM(1, "A_GAUSS") p->a_gus.push_back(v[0]); }
  for (int t = 0; t < c.ntrc; ++t) { NUM(1, "I_PHA") p->ipha.push_back((int)v[0]); }
  NUM(1, "N_FFT") c.nfft = (int)v[0];
  for (int t = 0; t < c.ntrc; ++t) { NEXT("OBS_FILES") p->obs_files.push_back(parse_string(line)); }
  NUM(2, "T_START T_END") c.t_start = v[0]; p->t_end = v[1];
  NUM(1, "DECONV_MODE") c.deconv_mode = (int)v[0];
  if (c.deconv_mode != 0 && c.deconv_mode != 1) { st = fail(RFINV_ERR_ARG, "ERROR: deconv_mode must be either 0 or 1%s", ""); goto done; }
  // SEA_DEP [BOREHOLE_DEP]: the second number is optional (src/params.f90:203-213, commented out in the reference)
  NEXT("SEA_DEP (BOREHOLE_DEP)")
  if (parse_numbers(line, 2, v)) { c.sdep = v[0]; c.bdep = v[1]; }
  else if (parse_numbers(line, 1, v)) { c.sdep = v[0]; c.bdep = 0.0; }
  else { st = fail(RFINV_ERR_IO, "ERROR: while reading SEA_DEP (BOREHOLE_DEP)%s", ""); goto done; }
  if (c.bdep < 0.0) { st = fail(RFINV_ERR_ARG, "ERROR: BOREHOLE_DEP must be positive%s", ""); goto done; }
  NEXT("VEL_FILE") p->vel_file = parse_string(line);
  NUM(1, "VP_MODE") c.vp_mode = (int)v[0];
  NUM(2, "K_MIN K_MAX") c.k_min = (int)v[0]; c.k_max = (int)v[1];
  NUM(2, "Z_MIN Z_MAX") c.z_min = v[0]; c.z_max = v[1];
  NUM(1, "H_MIN") c.h_min = v[0];
  NUM(1, "PRIOR_TYPE") c.prior_mode = (int)v[0];
  NUM(1, "DEV_DVS_PRIOR") c.dvs_prior = v[0];
  NUM(1, "DEV_DVP_PRIOR") c.dvp_prior = v[0];
  for (int t = 0; t < c.ntrc; ++t) { NUM(2, "SIG_MIN SIG_MAX") p->sig_min.push_back(v[0]); p->sig_max.push_back(v[1]); }
  NUM(1, "STEP_SIZE_Z") c.dev_z = v[0];
  NUM(1, "STEP_SIZE_DVS") c.dev_dvs = v[0];
  NUM(1, "STEP_SIZE_DVP") c.dev_dvp = v[0];
  NUM(1, "STEP_SIZE_SIG") c.dev_sig = v[0];
  NUM(1, "N_BIN_Z") c.nbin_z = (int)v[0];
  NUM(1, "N_BIN_VS") c.nbin_vs = (int)v[0];
  NUM(1, "N_BIN_VP") c.nbin_vp = (int)v[0];
  NUM(1, "N_BIN_VPVS") c.nbin_vpvs = (int)v[0];
  NUM(1, "N_BIN_SIG") c.nbin_sig = (int)v[0];
  NUM(1, "N_BIN_AMP") c.nbin_amp = (int)v[0];
  NUM(2, "AMP_MIN AMP_MAX") c.amp_min = v[0]; c.amp_max = v[1];
  NUM(2, "VP_MIN VP_MAX") c.vp_min = v[0]; c.vp_max = v[1];
  NUM(2, "VS_MIN VS_MAX") c.vs_min = v[0]; c.vs_max = v[1];
  NUM(2, "VPVS_MIN VPVS_MAX") c.vpvs_min = v[0]; c.vpvs_max = v[1];
#undef NUM
#undef NEXT
  {
    // observed traces; nsmp and delta are those of the last trace read (src/params.f90:446-449)
    int nsmp = 0;
    double delta = 0.0;
    std::vector<std::vector<double>> tr(c.ntrc);
    for (int t = 0; t < c.ntrc; ++t) {
      st = read_sac_window(join_path(p->base_dir, p->obs_files[t]), c.t_start, p->t_end, nsmp, delta, tr[t]);
      if (st != RFINV_OK) goto done;
    }
    c.nsmp = nsmp; c.delta = delta;
    p->obs.assign((size_t)c.ntrc * nsmp, 0.0);
    for (int t = 0; t < c.ntrc; ++t)
      for (int i = 0; i < nsmp && i < (int)tr[t].size(); ++i) p->obs[(size_t)t * nsmp + i] = tr[t][i];
    st = read_velmod(join_path(p->base_dir, p->vel_file), p);
    if (st != RFINV_OK) goto done;
  }
  c.rayps = p->rayps.data(); c.a_gus = p->a_gus.data(); c.ipha = p->ipha.data(); c.obs = p->obs.data(); c.r_inv = nullptr;
  c.vp_ref = p->vp_ref.data(); c.vs_ref = p->vs_ref.data(); c.sig_min = p->sig_min.data(); c.sig_max = p->sig_max.data();
done:
  if (st != RFINV_OK) { delete p; return st; }
  *out = p;
  return RFINV_OK;
}

void rfinv_problem_free(rfinv_problem* p) { delete p; }
const rfinv_config* rfinv_problem_config(const rfinv_problem* p) { return p ? &p->cfg : nullptr; }
const char* rfinv_problem_out_dir(const rfinv_problem* p) { return p ? p->out_dir.c_str() : ""; }
double rfinv_problem_t_end(const rfinv_problem* p) { return p ? p->t_end : 0.0; }

// Side outputs of get_params / read_obs: <out_dir>/params.in.copy and inputNN (src/params.f90:113-330, 462-468)
int32_t rfinv_problem_write_copies(const rfinv_problem* p, const char* out_dir, const char* input_dir) {
  if (!p) { rfinv_set_error("rfinv_problem_write_copies: NULL problem"); return RFINV_ERR_ARG; }
  const std::string od = out_dir && *out_dir ? out_dir : join_path(p->base_dir, p->out_dir);
  {
    std::ofstream f(od + "/params.in.copy");
    if (!f) return fail(RFINV_ERR_IO, "ERROR: cannot create %s/params.in.copy", od);
    for (const std::string& l : p->raw) f << " " << l << "\n";
  }
  const std::string idir = input_dir && *input_dir ? input_dir : (p->base_dir.empty() ? std::string(".") : p->base_dir);
  for (int t = 0; t < p->cfg.ntrc; ++t) {
    char name[32];
    snprintf(name, sizeof name, "input%02d", t + 1);
    std::ofstream f(idir + "/" + name);
    if (!f) return fail(RFINV_ERR_IO, "ERROR: cannot create %s", idir + "/" + name);
    for (int it = 0; it < p->cfg.nsmp; ++it)
      f << ld_real(it * p->cfg.delta + p->cfg.t_start) << ld_real(p->obs[(size_t)t * p->cfg.nsmp + it]) << "\n";
  }
  return RFINV_OK;
}

// output_results (src/mcmc_out.f90:97-318): the 12 files util/plot.py reads, from job-wide (already reduced) sums.
// Layouts: nk[k_max], nz[nbin_z], nsig[ntrc][nbin_sig], namp[ntrc][nsmp][nbin_amp], nvpz[nbin_vp][nbin_z],
// nvsz[nbin_vs][nbin_z], nvpvsz[nbin_vpvs][nbin_z], *_mean[nbin_z], likelihood_hist[nburn+niter] (sum over the cold
// chains of the job), vp_model/vs_model[n_models][nbin_z] (may be NULL / 0).
int32_t rfinv_write_outputs(const rfinv_config* c, const char* out_dir, int32_t nproc_total, int64_t nmod, const int64_t* nk,
                            const int64_t* nz, const int64_t* nsig, const int64_t* namp, const int64_t* nvpz,
                            const int64_t* nvsz, const int64_t* nvpvsz, const double* vp_mean, const double* vs_mean,
                            const double* vpvs_mean, const double* likelihood_hist, int32_t n_hist, const double* vp_model,
                            const double* vs_model, int64_t n_models) {
  if (!c || !out_dir || !nk || !nz || !nsig || !namp || !nvpz || !nvsz || !nvpvsz || !vp_mean || !vs_mean || !vpvs_mean) {
    rfinv_set_error("rfinv_write_outputs: NULL argument");
    return RFINV_ERR_ARG;
  }
  const std::string od = out_dir;
  const double dbin_amp = (c->amp_max - c->amp_min) / c->nbin_amp, dbin_vp = (c->vp_max - c->vp_min) / c->nbin_vp;
  const double dbin_vs = (c->vs_max - c->vs_min) / c->nbin_vs, dbin_z = (c->z_max - 0.0) / c->nbin_z;
  const double dbin_vpvs = (c->vpvs_max - c->vpvs_min) / c->nbin_vpvs;
  const double dn = (double)nmod;
  char buf[256];
  auto open = [&](const char* name, std::ofstream& f) -> int {
    f.open(od + "/" + name);
    if (!f) return fail(RFINV_ERR_IO, "ERROR: cannot create %s", od + "/" + name);
    return RFINV_OK;
  };
  int st;
  {  // all_models (src/mcmc_out.f90:110-131)
    std::ofstream f;
    if ((st = open("all_models", f)) != RFINV_OK) return st;
    for (int64_t im = 0; im < n_models && vp_model && vs_model; ++im) {
      if (vs_model[(size_t)im * c->nbin_z] < -900.0) continue;
      f << "\n";
      for (int iz = 1; iz <= c->nbin_z; ++iz)
        f << ld_real((iz - 0.5) * dbin_z) << ld_real(vp_model[(size_t)im * c->nbin_z + iz - 1])
          << ld_real(vs_model[(size_t)im * c->nbin_z + iz - 1]) << "\n";
      f << "\n";
    }
  }
  {  // likelihood (:134-144): mean logL of the cold chains per iteration
    std::ofstream f;
    if ((st = open("likelihood", f)) != RFINV_OK) return st;
    for (int it = 1; it <= n_hist && likelihood_hist; ++it)
      f << ld_int(it) << ld_real(likelihood_hist[it - 1] / (double)(c->ncool * nproc_total)) << "\n";
  }
  {  // num_interface.ppd (:147-158)
    std::ofstream f;
    if ((st = open("num_interface.ppd", f)) != RFINV_OK) return st;
    for (int ik = 1; ik <= c->k_max - 1; ++ik) f << ld_int(ik) << ld_real((double)nk[ik - 1] / dn) << "\n";
  }
  {  // syn_trace.ppd (:162-181) '(3F10.5,I6)'
    std::ofstream f;
    if ((st = open("syn_trace.ppd", f)) != RFINV_OK) return st;
    for (int t = 0; t < c->ntrc; ++t)
      for (int it = 1; it <= c->nsmp; ++it)
        for (int i = 1; i <= c->nbin_amp; ++i) {
          snprintf(buf, sizeof buf, "%10.5f%10.5f%10.5f%6d\n", (it - 1) * c->delta + c->t_start, c->amp_min + (i - 0.5) * dbin_amp,
                   (double)namp[((size_t)t * c->nsmp + it - 1) * c->nbin_amp + i - 1] / dn, t + 1);
          f << buf;
        }
  }
  {  // interface_depth.ppd (:184-196)
    std::ofstream f;
    if ((st = open("interface_depth.ppd", f)) != RFINV_OK) return st;
    for (int i = 1; i <= c->nbin_z; ++i) f << ld_real((i - 0.5) * dbin_z) << ld_real((double)nz[i - 1] / dn) << "\n";
  }
  {  // sigma.ppd (:199-217)
    std::ofstream f;
    if ((st = open("sigma.ppd", f)) != RFINV_OK) return st;
    for (int t = 0; t < c->ntrc; ++t) {
      if (!(c->sig_max[t] - c->sig_min[t] > (double)1.0e-5f)) continue;
      const double dbs = (c->sig_max[t] - c->sig_min[t]) / c->nbin_sig;
      for (int i = 1; i <= c->nbin_sig; ++i)
        f << ld_real((i - 0.5) * dbs + c->sig_min[t]) << ld_real((double)nsig[(size_t)t * c->nbin_sig + i - 1] / dn) << ld_int(t + 1) << "\n";
    }
  }
  struct Prof { const char* name; const int64_t* h; int nb; double db, vmin; };
  const Prof profs[3] = {{"vs_z.ppd", nvsz, c->nbin_vs, dbin_vs, c->vs_min}, {"vp_z.ppd", nvpz, c->nbin_vp, dbin_vp, c->vp_min},
                         {"vpvs_z.ppd", nvpvsz, c->nbin_vpvs, dbin_vpvs, c->vpvs_min}};
  for (const Prof& pr : profs) {  // (:220-273) '(3F10.5)'
    std::ofstream f;
    if ((st = open(pr.name, f)) != RFINV_OK) return st;
    for (int iv = 1; iv <= pr.nb; ++iv)
      for (int iz = 1; iz <= c->nbin_z; ++iz) {
        snprintf(buf, sizeof buf, "%10.5f%10.5f%10.5f\n", (iv - 0.5) * pr.db + pr.vmin, (iz - 0.5) * dbin_z,
                 (double)pr.h[(size_t)(iv - 1) * c->nbin_z + iz - 1] / dn);
        f << buf;
      }
  }
  struct Mean { const char* name; const double* m; };
  const Mean means[3] = {{"vs_z.mean", vs_mean}, {"vp_z.mean", vp_mean}, {"vpvs_z.mean", vpvs_mean}};
  for (const Mean& mn : means) {  // (:276-317) '(2F10.5)'
    std::ofstream f;
    if ((st = open(mn.name, f)) != RFINV_OK) return st;
    for (int iz = 1; iz <= c->nbin_z; ++iz) {
      snprintf(buf, sizeof buf, "%10.5f%10.5f\n", mn.m[iz - 1] / dn, (iz - 0.5) * dbin_z);
      f << buf;
    }
  }
  return RFINV_OK;
}

}  // extern "C"
