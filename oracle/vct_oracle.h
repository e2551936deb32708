/*
 * vct_oracle.h -- C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a scalar restatement of the reference's GLSL passes and of the GL fixed-function
 * rules they depend on.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The shipped library
 * (libvct_b200.so) never links or calls anything in this directory.
 *
 * PARITY: the reference (AlerianEmperor/Voxel-Cone-Tracing) has no tests, golden vectors or
 * fixtures, cannot be compiled in this image (Windows-only source, glm/assimp/GL absent) and never
 * reads anything back.  Its shader files CAN be executed: tests/glsl_run.py interprets them and
 * tests/golden/reference_shader_vectors.npz holds their outputs (generator committed next to it);
 * the oracle reproduces those byte for byte (tests/test_reference_glsl.py).  The fixed-function GL
 * stages between the shaders remain UNPINNED by the reference (GL 4.3 specification restated).
 * Further pins: the hand-derivable known-answer vectors of SURVEY.md A.7 (tests/test_oracle_kat.py),
 * tests/test_oracle_independent.py and the fixtures under tests/golden/.
 */
#ifndef VCT_ORACLE_H_
#define VCT_ORACLE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_CONES 16

/* Every field mirrors a uniform / constant of the reference; see DESIGN.md "Parameters". */
typedef struct orc_params {
  int32_t VoxelDimensions;           /* Voxel_Cone_Tracing.h:16  (128) */
  float   VoxelGridWorldSize;        /* Voxel_Cone_Tracing.h:17  (150) */
  int32_t ShadowMapSize;             /* Voxel_Cone_Tracing.h:35  (4096) */
  int32_t screen_width;              /* Voxel_Cone_Tracing.h:24 */
  int32_t screen_height;             /* Voxel_Cone_Tracing.h:25 */
  float   ModelMatrix[16];           /* column-major, Voxel_Cone_Tracing.h:183 */
  float   ModelViewMatrix[16];       /* Voxel_Cone_Tracing.h:185 */
  float   ProjectionMatrix[16];      /* Voxel_Cone_Tracing.h:186 */
  float   DepthModelViewProjectionMatrix[16]; /* Voxel_Cone_Tracing.h:187,205,241 */
  float   ProjX[16], ProjY[16], ProjZ[16];    /* Voxel_Cone_Tracing.h:130-134 */
  float   CameraPosition[3];         /* Voxel_Cone_Tracing.h:167 */
  float   LightDirection[3];         /* Voxel_Cone_Tracing.h:168 */
  float   ambientFactor;             /* Voxel_Cone_Tracing.h:171 */
  int32_t NumDiffuseCones;           /* VoxelConeTracing.fs:46 */
  float   ConeDirections[ORC_MAX_CONES * 3]; /* VoxelConeTracing.fs:49-57 */
  float   ConeWeights[ORC_MAX_CONES];        /* VoxelConeTracing.fs:48 */
  float   DiffuseTanHalfAngle;       /* VoxelConeTracing.fs:198 (0.577) */
  float   SpecularTanHalfAngle;      /* VoxelConeTracing.fs:218 (0.07) */
  float   StepMultiplier;            /* implicit 1.0, VoxelConeTracing.fs:103 */
  float   MaxDistance;               /* VoxelConeTracing.fs:43 (75) */
  float   MaxAlpha;                  /* VoxelConeTracing.fs:44 (0.95) */
  int32_t PcfRadius;                 /* Voxelization.fs:26 (2) */
  float   ShadowBias;                /* Voxelization.fs:88 (0.002) */
  int32_t CoveragePolicy;            /* 0 CENTER, 1 MSAA4_ANY, 2 CONSERVATIVE */
  int32_t VoxelStoreMode;            /* 0 sum+count average, 1 last writer in primitive order */
  int32_t Bounces;                   /* 2 = reference; >=3 = voxel-space re-injection extension */
  int32_t FilterMode;                /* voxel-texture filter weights: 1 (default) = 8 fractional bits, rounded (LOD
                                        fraction truncated), as measured on B200 texture hardware and as llvmpipe's
                                        RGBA8 path does; 0 = fp32 weights */
  int32_t GridFormat;                /* 0 = RGBA8 (the reference, Voxel_Cone_Tracing.h:119), 1 = RGBA16F (BASELINE config 3) */
  int32_t RasterOrigin;              /* frame pass edge functions: 0 = window origin (defined semantics, what the CUDA path
                                        does), 1 = per-triangle local origin (oracle-only experiment) */
} orc_params;

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(void);
void     orc_destroy(orc_ctx*);
void     orc_default_params(orc_params* p);
int      orc_set_params(orc_ctx*, const orc_params* p);

/* channels in {1,3,4}; expanded to RGBA8 following Model.h:159-169 + GL swizzle rules. */
int orc_upload_texture(orc_ctx*, int id, int w, int h, int channels, const uint8_t* pix);
int orc_set_material(orc_ctx*, int mat, int diffuse, int specular, int height, float shininess);
/* verts14: Mesh.h:12-19 (pos3 nrm3 uv2 tan3 bitan3); idx: 3 per triangle; tri_material may be NULL (all 0) */
int orc_upload_mesh(orc_ctx*, const float* verts14, size_t nv, const uint32_t* idx, size_t nt,
                    const uint16_t* tri_material);

int orc_draw_depth(orc_ctx*);          /* S1 */
int orc_draw_voxels(orc_ctx*);         /* clear + V1..V4 + resolve + M1 (+ re-injection when Bounces>=3) */
int orc_draw_voxels_range(orc_ctx*, size_t tri_begin, size_t tri_end, int clear_first); /* accumulate only */
int orc_resolve_and_mip(orc_ctx*);
int orc_render(orc_ctx*);              /* S2 + C1..C6 */
int orc_render_rows(orc_ctx*, int y_begin, int y_end);   /* same, rows [y_begin, y_end) only */

int orc_get_depth(orc_ctx*, uint32_t* d24);                 /* S*S */
int orc_get_counts(orc_ctx*, uint32_t* counts);             /* V^3, index (z*V+y)*V+x */
int orc_get_sums(orc_ctx*, uint32_t* rgb_sums);             /* V^3*3 */
int orc_set_accum(orc_ctx*, const uint32_t* counts, const uint32_t* rgb_sums);
int orc_get_grid(orc_ctx*, int level, uint8_t* rgba);       /* (V>>level)^3*4 bytes (RGBA8) or *8 (RGBA16F half bits) */
int orc_set_grid_level0(orc_ctx*, const uint8_t* rgba);     /* then orc_build_mips */
int orc_build_mips(orc_ctx*);
int orc_get_visibility(orc_ctx*, uint32_t* tri_id);         /* H*W, 0xFFFFFFFF = background */
int orc_get_frame(orc_ctx*, uint8_t* rgba);                 /* H*W*4, row 0 = bottom (GL window coords) */
uint64_t orc_cone_samples(orc_ctx*);
uint64_t orc_fragment_count(orc_ctx*);                      /* voxel fragments of the last draw_voxels */
/* OpenMP threads of the oracle's parallel loops: n > 0 sets them (torchrun exports OMP_NUM_THREADS=1, which would
 * silently turn the CPU baseline into a one-thread run); returns the number in effect. */
int orc_set_num_threads(int n);

/* point probes for known-answer tests */
void orc_sample_voxels(orc_ctx*, const float world_pos[3], float lod, float out_rgba[4]);
void orc_cone(orc_ctx*, const float start[3], const float dir[3], float tan_half, float out_rgba[4],
              int* n_steps);
int  orc_select_axis(const float w0[3], const float w1[3], const float w2[3]);
void orc_sample_texture(orc_ctx*, int tex, float u, float v, float lod, float out_rgba[4]);
float orc_pcf(orc_ctx*, const float dc[4]);                 /* returns lit tap count / taps (V3 normalisation) */

#ifdef __cplusplus
}
#endif
#endif
