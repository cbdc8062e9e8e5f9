// Internal definition of the opaque rfinv_handle.
#pragma once
#include <vector>
#include "rfinv_common.cuh"

struct PtState;  // pt.cu

struct rfinv_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  rfinv_config cfg;  // deep copy; pointers refer to the vectors below
  std::vector<double> h_rayps, h_a_gus, h_obs, h_vp_ref, h_vs_ref, h_sig_min, h_sig_max, h_r_inv;
  std::vector<int32_t> h_ipha;
  DevConfig dc;
  // constants in HBM
  double* d_flt = nullptr;
  double2* d_tw = nullptr;
  double* d_obs = nullptr;
  double* d_vp_ref = nullptr;
  double* d_vs_ref = nullptr;
  double* d_r_inv = nullptr;
  double* d_w_fac = nullptr;
  // evaluation workspace (grown on demand)
  int cap = 0;
  int* d_k = nullptr;
  double *d_z = nullptr, *d_dvp = nullptr, *d_dvs = nullptr, *d_sig = nullptr, *d_stage = nullptr;
  double *d_misfit = nullptr, *d_phi = nullptr, *d_logl = nullptr, *d_rft_full = nullptr, *d_scratch = nullptr;
  uint8_t* d_valid = nullptr;
  double* d_qpart = nullptr;    // quadform_kernel partial sums
  int* d_qcnt = nullptr;        // quadform_kernel work / arrival counters
  // rfinv_eval_batch: the models go up as the caller holds them (chain slowest, no layout kernels); large batches in
  // pieces on a second stream while prep_kernel already works on the pieces that have landed
  cudaStream_t stream_copy = nullptr;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr};
  int* d_ready = nullptr;       // [pieces] + 1: epoch of the last upload that filled the piece | time-out flag
  int* h_ready = nullptr;       // pinned: [0] the epoch being written, [1] time-out flag read back
  int ready_cap = 0, ready_epoch = 0;
  size_t cap_rft_full = 0;
  int launches = 0;
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // parallel-tempering state (pt.cu)
  PtState* pt = nullptr;

  int ensure_capacity(int C);
  void free_workspace();
  void free_pt();
  // forward + quadratic form (+ logL) for device-resident chain-fastest arrays
  int eval_device(int C, const int* k, const double* z, const double* dvp, const double* dvs, const double* sig,
                  double* logl, double* rft_smp, double* rft_full, uint8_t* is_valid, const int* active, int n_active,
                  const ModelBatch* layout = nullptr);   // layout: chain_major / ready* fields to take over (host path)
};
