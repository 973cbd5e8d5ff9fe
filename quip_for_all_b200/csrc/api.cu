// C-ABI glue: version / error strings / device info / tuning options.
#include <string.h>

#include "common.cuh"

namespace qb {
int64_t g_launch_count = 0;
int g_opt_pdl = 1;   // programmatic dependent launch between our kernels
extern int g_opt_table_repl;
extern int g_opt_rot_cluster;
extern int g_opt_gemv_warps;
extern int g_opt_gemv_ctas_per_sm;
extern int g_opt_stage_mask;
extern int g_opt_fuse;
extern int g_ds_flags;
extern int g_opt_phase0;
extern int g_opt_lean;
extern int g_opt_epi_mma;
extern int g_opt_nearest_struct;
extern int g_opt_rot_warp_rows;
extern int g_opt_rot_pipe_rows;
int g_opt_umma_rt = 0;       // experiments: 1 = one 128-row tile per CTA at every M (0 = auto: two when M > 64)
int g_opt_umma_ksplit = 0;   // experiments: force the split-K factor of the tcgen05 kernel (0 = auto)
int g_opt_umma = 2;     // in-kernel decode + tcgen05 GEMM (umma_gemm.cu): 0 = never, 1 = whenever the shape is covered
                        // (M <= 256), 2 = auto: where it measured faster than the alternatives (profiles/README.md)
}  // namespace qb

extern "C" int quipb200_abi_version(void) { return QUIPB200_ABI_VERSION; }

extern "C" const char* quipb200_strerror(int code) {
  if (code == 0) return "success";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case QUIPB200_EINVAL: return "quipb200: invalid argument (shape / null pointer / unsupported combination)";
    case QUIPB200_EALIGN: return "quipb200: pointer or row pitch is not 16-byte aligned";
    case QUIPB200_EWORKSPACE: return "quipb200: workspace too small";
    case QUIPB200_EUNSUPPORTED: return "quipb200: shape not covered by the fused path (use the dense path)";
  }
  return "quipb200: unknown error";
}

extern "C" int quipb200_sm_count(void) {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" int quipb200_set_option(const char* name, int value) {
  if (!name) return QUIPB200_EINVAL;
  if (!strcmp(name, "gemv_table_repl")) {
    if (value != 1 && value != 16) return QUIPB200_EINVAL;
    qb::g_opt_table_repl = value;
    return 0;
  }
  if (!strcmp(name, "rot_cluster")) {
    qb::g_opt_rot_cluster = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "gemv_warps")) {
    if (value < 0 || value > 32) return QUIPB200_EINVAL;
    qb::g_opt_gemv_warps = value;
    return 0;
  }
  if (!strcmp(name, "gemv_ctas_per_sm")) {
    if (value < 1 || value > 4) return QUIPB200_EINVAL;
    qb::g_opt_gemv_ctas_per_sm = value;
    return 0;
  }
  if (!strcmp(name, "stage_mask")) {
    if (value < 0 || value > 7) return QUIPB200_EINVAL;
    qb::g_opt_stage_mask = value;
    return 0;
  }
  if (!strcmp(name, "lean")) {
    qb::g_opt_lean = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "phase0")) {
    qb::g_opt_phase0 = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "epi_mma")) {
    qb::g_opt_epi_mma = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "nearest_struct")) {
    qb::g_opt_nearest_struct = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "pdl")) {
    qb::g_opt_pdl = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(name, "umma_rt")) {
    qb::g_opt_umma_rt = value;
    return 0;
  }
  if (!strcmp(name, "umma_ksplit")) {
    if (value < 0 || value > 16) return QUIPB200_EINVAL;
    qb::g_opt_umma_ksplit = value;
    return 0;
  }
  if (!strcmp(name, "umma")) {
    if (value < 0 || value > 2) return QUIPB200_EINVAL;
    qb::g_opt_umma = value;
    return 0;
  }
  if (!strcmp(name, "rot_pipe_rows")) {
    if (value < 1) return QUIPB200_EINVAL;
    qb::g_opt_rot_pipe_rows = value;
    return 0;
  }
  if (!strcmp(name, "rot_warp_rows")) {
    if (value < 1) return QUIPB200_EINVAL;
    qb::g_opt_rot_warp_rows = value;
    return 0;
  }
  if (!strcmp(name, "ds_flags")) {
    qb::g_ds_flags = value;
    return 0;
  }
  if (!strcmp(name, "fuse")) {
    if (value < 0 || value > 3) return QUIPB200_EINVAL;
    qb::g_opt_fuse = value;
    return 0;
  }
  return QUIPB200_EINVAL;
}

extern "C" int quipb200_get_option(const char* name) {
  if (!name) return QUIPB200_EINVAL;
  if (!strcmp(name, "gemv_table_repl")) return qb::g_opt_table_repl;
  if (!strcmp(name, "rot_cluster")) return qb::g_opt_rot_cluster;
  if (!strcmp(name, "epi_mma")) return qb::g_opt_epi_mma;
  if (!strcmp(name, "nearest_struct")) return qb::g_opt_nearest_struct;
  if (!strcmp(name, "gemv_warps")) return qb::g_opt_gemv_warps;
  if (!strcmp(name, "gemv_ctas_per_sm")) return qb::g_opt_gemv_ctas_per_sm;
  if (!strcmp(name, "stage_mask")) return qb::g_opt_stage_mask;
  if (!strcmp(name, "fuse")) return qb::g_opt_fuse;
  if (!strcmp(name, "ds_flags")) return qb::g_ds_flags;
  if (!strcmp(name, "pdl")) return qb::g_opt_pdl;
  if (!strcmp(name, "umma_ksplit")) return qb::g_opt_umma_ksplit;
  if (!strcmp(name, "umma")) return qb::g_opt_umma;
  if (!strcmp(name, "rot_warp_rows")) return qb::g_opt_rot_warp_rows;
  if (!strcmp(name, "rot_pipe_rows")) return qb::g_opt_rot_pipe_rows;
  return QUIPB200_EINVAL;
}

extern "C" int64_t quipb200_launch_count(void) { return qb::g_launch_count; }
