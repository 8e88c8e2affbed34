/* plenoctree_b200.h — C ABI of the B200-native NeRF-SH hot path.
 *
 * The reference (sxyu/plenoctree) has no FFI layer of its own: its boundary for this path is a
 * set of Python call signatures.  Each entry point below is what a binding for
 * one of those call sites would call; the reference interface it replaces is cited as
 * file:line relative to the reference tree.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - every pointer named *_dev is a CUDA device pointer on the current device, *_host is host
 *     memory; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - all functions return 0 on success, non-zero on error; pob_last_error() returns a
 *     thread-local message for the last failure (never NULL);
 *   - nothing here falls back to the CPU: without a CUDA device every compute call fails.
 *   - precision: POB_PREC_FP16 = fp16 operands / fp32 accumulate on tcgen05 (the numerics class
 *     of the reference's TF32-default XLA GPU path); POB_PREC_FP16X3 = error-compensated 3-pass
 *     split (fp32-class accuracy, 1/3 of the tensor throughput).
 */
#ifndef PLENOCTREE_B200_H
#define PLENOCTREE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POB_PREC_FP16 1
#define POB_PREC_FP16X3 3

/* ---------------------------------------------------------------------------------------------
 * Library / device
 * ------------------------------------------------------------------------------------------- */
int pob_abi_version(void);
const char* pob_last_error(void);
/* number of SMs of the current CUDA device (persistent-grid size); <0 on error */
int pob_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Parameters of one MLP (MLP_0 coarse / MLP_1 fine; nerf_sh/nerf/models.py:83-104).
 * Flat fp32 layout = Dense_0..Dense_9 in order, each kernel [in,out] row-major then bias [out]
 * (nerf_sh/nerf/model_utils.py:60-93; Dense index mapping octree/nerf/models.py:79-102).
 * ------------------------------------------------------------------------------------------- */
/* number of fp32 parameters of one MLP for SH degree sh_deg (-1 = plain RGB head, 3 channels) */
int64_t pob_param_count(int sh_deg);
/* byte size of the packed tensor-core operand blob of one MLP */
int64_t pob_packed_bytes(int sh_deg);
/* flat fp32 parameters -> packed blob (fp16 hi/lo forward images, transposed images, biases) */
int pob_pack_weights(const float* flat_dev, int sh_deg, void* packed_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * NerfModel.eval_points_raw(points, viewdirs=None, coarse=False) -> (raw_rgb[M,3K], raw_sigma[M,1])
 *   JAX:   nerf_sh/nerf/models.py:143-181      torch twin: octree/nerf/models.py:211-252
 * raw_rgb_dev may be NULL (sigma only: octree/extraction.py:271,316 discard rgb).
 * ------------------------------------------------------------------------------------------- */
int pob_eval_points_raw(const void* packed_dev, int sh_deg, const float* points_dev, int64_t m,
                        float* raw_rgb_dev, float* raw_sigma_dev, int precision, void* stream);

/* NerfModel.eval_points(points, viewdirs) -> (rgb[M,3], sigma[M,1])  (models.py:183-214):
 * eval_sh at the per-point view direction, sigmoid / relu.  out_rgbs_dev: [M,4] = (r,g,b,sigma). */
int pob_eval_points(const void* packed_dev, int sh_deg, const float* points_dev,
                    const float* viewdirs_dev, int64_t m, float* out_rgbs_dev, int precision,
                    void* stream);

/* Dense-grid sweep of octree.extraction (auto_scale / step1: octree/extraction.py:244-320).
 * Evaluates raw sigma (and optionally raw SH coefficients) at the voxel centres
 *   ((i + 0.5)/reso - offset[a]) / scale[a]
 * of the slab ix in [x0, x0+nx), iy in [0,ny), iz in [0,nz), flattened x-major like
 * torch.meshgrid(xx,yy,zz).reshape(3,-1).T.  No host grid, no H2D copies. */
int pob_eval_grid(const void* packed_dev, int sh_deg, int reso, int x0, int nx, int ny, int nz,
                  const float offset[3], const float scale[3], float* raw_rgb_dev,
                  float* raw_sigma_dev, int precision, void* stream);

/* Host-buffer convenience form of pob_eval_points_raw (H2D, kernel, D2H inside the call);
 * the e2e arm of bench.py times this. */
int pob_eval_points_raw_host(const void* packed_dev, int sh_deg, const float* points_host,
                             int64_t m, float* raw_rgb_host, float* raw_sigma_host,
                             int precision);

/* ---------------------------------------------------------------------------------------------
 * Test bench for the tcgen05 descriptor conventions (tests/test_umma_probe.py).
 * Runs `nops` tcgen05.mma (kind::f16, M=128) on two shared-memory images and returns the
 * [128 x out_cols] fp32 accumulator.
 * ------------------------------------------------------------------------------------------- */
int pob_umma_probe(const void* a_img_dev, uint32_t a_bytes, const void* b_img_dev,
                   uint32_t b_bytes, uint32_t b_off, const uint64_t* adesc_dev,
                   const uint64_t* bdesc_dev, const uint32_t* dcol_dev, const uint32_t* accum_dev,
                   int nops, uint32_t idesc, int out_cols, float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLENOCTREE_B200_H */
