/* b200fem.h -- C ABI of the B200-native matrix-free operator-apply path.
 *
 * Drop-in boundary for sfem's matrix-free hot path. Every entry point is
 * extern "C", takes plain pointers and sizes, returns int (0 = B200FEM_SUCCESS, 1 = B200FEM_FAILURE,
 * same convention as SFEM_SUCCESS / SFEM_FAILURE, base/sfem_base.hpp:12-13 of the reference) and
 * never throws. There is NO CPU fallback: a call without a usable CUDA device fails.
 *
 * Conventions shared with the reference's operator-level C functions
 * (operators/laplacian.hpp, operators/linear_elasticity.hpp, operators/.../cuda/cu_*.cu):
 *   - connectivity is SoA `elements[nxe][nelements]`, int32 (idx_t), HOST memory at plan creation;
 *   - geometry factors are fp32 (jacobian_t / geom_t) even for fp64 vectors, HOST memory at plan
 *     creation: FFF `[nelements*6]` AoS, adjugate `[nelements*9]` AoS, determinant `[nelements]`;
 *   - vectors are DEVICE pointers of `real_type` (B200FEM_F64 / B200FEM_F32);
 *   - vector fields (linear elasticity) are AoS `[nnodes*3]` = the reference's block_size-3 layout
 *     (frontend/ops/sfem_LinearElasticity.cpp:540-550 hands &u[0],&u[1],&u[2] with stride 3);
 *   - apply ACCUMULATES: y += A x (callers zero first: algebra/sfem_cg.hpp:226-230);
 *   - `stream` is a cudaStream_t passed by value as void* (NULL = default stream); all launches are
 *     asynchronous with respect to the host unless stated otherwise.
 */
#ifndef B200FEM_H
#define B200FEM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FEM_SUCCESS 0
#define B200FEM_FAILURE 1

/* real_t of the vectors; value == sizeof */
#define B200FEM_F32 4
#define B200FEM_F64 8

/* element families (numeric values follow smesh::ElemType as far as the reference shows them) */
#define B200FEM_TET4 4
#define B200FEM_HEX8 8
#define B200FEM_SSHEX8 800 /* semi-structured hex8; the refinement level is a separate argument */
#define B200FEM_TET10 10        /* quadratic tetrahedron; nodes 0-3 corners, 4 = edge(0,1), 5 = (1,2), 6 = (0,2), 7 = (0,3),
                                   8 = (1,3), 9 = (2,3) (operators/tet10/tet10_laplacian_inline_cpu.hpp:6-28)               */
#define B200FEM_MACRO_TET4 1004 /* P1-iso-P2: the same 10 nodes as 8 linear sub-tets
                                   (operators/macro_tet4/macro_tet4_laplacian.cpp:12-125)                                     */

/* scatter strategies of b200fem_op_set_option("scatter", ...) */
#define B200FEM_SCATTER_AUTO 0
#define B200FEM_SCATTER_ATOMIC 1  /* red.global.add on every element dof                        */
#define B200FEM_SCATTER_COLORED 2 /* element colouring, plain read-modify-write, one launch/colour */
#define B200FEM_SCATTER_PATCH 3   /* element patches staged in shared memory, fixed-order sums, one red per (patch, node):
                                     the default of tet4 Laplacian, hex8 and tet10 linear elasticity (hex8 / tet10 / macro_tet4
                                     Laplacian and macro_tet4 linear elasticity: ATOMIC -- each default is the measured faster one) */

typedef struct b200fem_op_s *b200fem_op_t; /* opaque operator plan (device connectivity + geometry) */

/* ------------------------------------------------------------------------------------------
 * Runtime helpers (what smesh::Buffer / sfem::BLAS<T>::allocate give the reference's device path,
 * algebra/cuda/sfem_cuda_blas.cu:497-514)
 * ---------------------------------------------------------------------------------------- */
const char *b200fem_last_error(void);          /* message of the last failure on this thread        */
int b200fem_device_count(int *count);
int b200fem_set_device(int device);
int b200fem_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *global_mem_bytes);
int b200fem_malloc(void **ptr, size_t bytes);  /* zero-filled, like BLAS<T>::allocate (calloc)      */
int b200fem_free(void *ptr);
int b200fem_host_malloc(void **ptr, size_t bytes); /* pinned host memory                            */
int b200fem_host_free(void *ptr);
int b200fem_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int b200fem_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int b200fem_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int b200fem_stream_create(void **stream);
int b200fem_stream_destroy(void *stream);
int b200fem_stream_synchronize(void *stream);
int b200fem_device_synchronize(void);

/* ------------------------------------------------------------------------------------------
 * Geometry caches (setup path). Replace tet4_fff_fill (operators/tet4/tet4_fff.cpp:4-27),
 * hex8_fff_fill (operators/hex8/hex8_fff.cpp:5-37) and hex8_adjugate_and_det_fill
 * (operators/hex8/hex8_jacobian.cpp:5-38; TET4: tet4_adjugate_and_det, operators/tet4/tet4_inline_cpu.hpp:285-380).
 * Host in, host out; the work runs on the device.
 * For B200FEM_SSHEX8 pass the 8 macro corners as `elements` (rows of the level-L lattice corners).
 * ---------------------------------------------------------------------------------------- */
int b200fem_fff_fill(int element_type, ptrdiff_t nelements, ptrdiff_t nnodes, const int32_t *const *elements,
                     const float *const *points, float *fff /* [nelements*6] */);
int b200fem_adjugate_and_det_fill(int element_type, ptrdiff_t nelements, ptrdiff_t nnodes,
                                  const int32_t *const *elements, const float *const *points,
                                  float *adjugate /* [nelements*9] */, float *determinant /* [nelements] */);

/* ------------------------------------------------------------------------------------------
 * Operator plans.
 *   b200fem_laplacian_create   <- laplacian_apply_opt(type, ne, elements, fff, u, out)
 *                                 (operators/laplacian.cpp:197-236): TET4 -> tet4_laplacian_apply_opt
 *                                 (operators/tet4/tet4_laplacian.cpp:237), HEX8 -> hex8_laplacian_apply_opt
 *                                 (operators/hex8/hex8_laplacian.cpp:309), SSHEX8 ->
 *                                 affine_sshex8_laplacian_stencil_apply_fff
 *                                 (operators/sshex8/sshex8_laplacian.cpp:698), TET10 -> tet10_laplacian_apply_opt /
 *                                 _diag_opt (operators/tet10/tet10_laplacian.cpp:232-302), MACRO_TET4 ->
 *                                 macro_tet4_laplacian_apply_opt / _diag_opt
 *                                 (operators/macro_tet4/macro_tet4_laplacian.cpp:401-490); the 10-node families take the
 *                                 tet4 FFF of their 4 corners. tet10 `diag` reproduces the reference's value, which is
 *                                 1/6 of the diagonal of its own apply (option "reference_diag" = 0 gives the consistent
 *                                 one). GPU twin being superseded:
 *                                 cu_laplacian_apply (frontend/cuda/sfem_Function_incore_cuda.cpp:683-697).
 *   b200fem_linear_elasticity_create <- TET4: tet4_linear_elasticity_apply_opt (operators/tet4/tet4_linear_elasticity.cpp:237-300),
 *                                 HEX8: affine_hex8_linear_elasticity_apply_adjugate_soa
 *                                 (operators/hex8/hex8_linear_elasticity.cpp:625), SSHEX8:
 *                                 affine_sshex8_linear_elasticity_apply_macro_adjugate
 *                                 (operators/sshex8/sshex8_linear_elasticity.cpp:275), TET10:
 *                                 tet10_linear_elasticity_apply_opt / _diag_opt
 *                                 (operators/tet10/tet10_linear_elasticity.cpp:209-355), MACRO_TET4:
 *                                 macro_tet4_linear_elasticity_apply_opt / _diag_opt
 *                                 (operators/macro_tet4/macro_tet4_linear_elasticity.cpp:206-277, 397-437); both with the
 *                                 tet4 adjugate / determinant of the 4 corners.
 * `level` is ignored unless element_type == B200FEM_SSHEX8 (then nxe = (level+1)^3 rows of
 * `elements`, local lattice index z*(L+1)^2 + y*(L+1) + x, operators/sshex8/sshex8.hpp:18).
 * The plan copies and re-lays-out connectivity and geometry on the current device; the host
 * arrays may be released after the call. No allocation happens in apply.
 * ---------------------------------------------------------------------------------------- */
int b200fem_laplacian_create(b200fem_op_t *op, int element_type, int level, ptrdiff_t nelements, ptrdiff_t nnodes,
                             const int32_t *const *elements, const float *fff, int real_type);
int b200fem_linear_elasticity_create(b200fem_op_t *op, int element_type, int level, ptrdiff_t nelements,
                                     ptrdiff_t nnodes, const int32_t *const *elements, const float *adjugate,
                                     const float *determinant, double mu, double lambda, int real_type);
/* The reference's non-default paths that take the node coordinates (fp32 SoA `points[3][nnodes]`, HOST) instead of cached
 * geometry factors:
 *   laplacian, TET4   <- tet4_laplacian_apply (operators/tet4/tet4_laplacian.cpp:70-135): FFF recomputed per element from
 *                        the points in the vector precision (used when the Op has no FFF cache, sfem_Laplacian.cpp:43-49);
 *   laplacian, SSHEX8 <- sshex8_laplacian_apply (operators/sshex8/sshex8_laplacian.cpp:39-182): trilinear (non-affine) macro
 *                        elements, FFF at the 6 face-centre quadrature points of every micro element; `points` is indexed by
 *                        sshex8 node id, only the 8 macro corners of each element are read;
 *   linear elasticity, TET4 <- tet4_linear_elasticity_apply / _diag / _block_diag_sym (operators/tet4/tet4_linear_elasticity.cpp:
 *                        86-230, 360-440; the reference's block diagonal is 6 x the blocks of its apply: reproduced, option
 *                        "reference_block_diag" = 0 gives the consistent ones; also on the cached-adjugate TET4 plan): the
 *                        reference's only tet4 elasticity path (its cached-adjugate twin is declared but not defined);
 *   linear elasticity, HEX8 <- hex8_linear_elasticity_apply (operators/hex8/hex8_linear_elasticity.cpp:86-324): the
 *                        isoparametric variant of SFEM_HEX8_ASSUME_AFFINE=0, Jacobian at each of the 2x2x2 Gauss points;
 *   laplacian / linear elasticity, TET10 and MACRO_TET4 <- tet10_laplacian_apply / _diag (operators/tet10/tet10_laplacian.cpp:
 *                        65-230), macro_tet4_laplacian_apply / _diag (operators/macro_tet4/macro_tet4_laplacian.cpp:127-260),
 *                        tet10_linear_elasticity_apply / _diag (operators/tet10/tet10_linear_elasticity.cpp:21-207),
 *                        macro_tet4_linear_elasticity_apply / _diag (operators/macro_tet4/macro_tet4_linear_elasticity.cpp:
 *                        137-204, 330-395): geometry of the 4 corners recomputed per element in the vector precision. */
int b200fem_laplacian_create_from_points(b200fem_op_t *op, int element_type, int level, ptrdiff_t nelements, ptrdiff_t nnodes,
                                         const int32_t *const *elements, const float *const *points, int real_type);
int b200fem_linear_elasticity_create_from_points(b200fem_op_t *op, int element_type, int level, ptrdiff_t nelements,
                                                 ptrdiff_t nnodes, const int32_t *const *elements, const float *const *points,
                                                 double mu, double lambda, int real_type);
/* Mass and LumpedMass, the neighbouring operators on the same gather -> local -> scatter skeleton:
 *   apply / apply_overwrite <- Mass::apply -> apply_mass (frontend/ops/sfem_Mass.cpp:107-125, operators/mass.cpp:158-207) ->
 *                        hex8_apply_mass (operators/hex8/hex8_mass.cpp:8-60): isoparametric Q1, 27-point Gauss. TET4 fails,
 *                        as in the reference (its dispatcher has no tet4 case);
 *   diag               <- LumpedMass::hessian_diag -> assemble_lumped_mass (frontend/ops/sfem_LumpedMass.cpp, operators/mass.cpp:
 *                        65-114) -> hex8_assemble_lumped_mass (hex8_mass.cpp:62-103) / tet4_assemble_lumped_mass
 *                        (operators/tet4/tet4_mass.cpp:231-282) / tet10_assemble_lumped_mass (operators/tet10/tet10_mass.cpp:
 *                        443-583; TET10, 10 rows of `elements`: -det/120 on corners, det/30 on edge nodes, as the reference):
 *                        diag += row sums. Scalar fields (block_size 1). */
/* Packed meshes: the Laplacian on smesh::PackedMesh<uint16_t> connectivity, the data PackedLaplacian::apply hands to
 * packed_laplacian_apply<uint16_t> (frontend/ops/sfem_PackedLaplacian.cpp:111-176, 551-612, 875-930; FunctionSpace::PackedIdxType,
 * frontend/sfem_FunctionSpace.hpp:19). Elements are cut into packs of n_elements_per_pack consecutive elements; pack p owns the
 * contiguous node range [owned_nodes_ptr[p], owned_nodes_ptr[p+1]), the last n_shared[p] of which other packs touch too; its
 * other nodes are ghost_idx[ghost_ptr[p] .. ghost_ptr[p+1]); `elements[v][e]` is LOCAL to the pack of e: l < n_owned is node
 * owned_nodes_ptr[p] + l, else ghost l - n_owned. TET4, HEX8, TET10 (the reference's three micro kernels) and MACRO_TET4;
 * fff [nelements * 6] as for b200fem_laplacian_create. All arrays HOST, copied by the plan. Only apply / apply_overwrite exist
 * (the reference's PackedLaplacian implements nothing else). Option "packed_kernel": 1 = direct kernel (default), 0 = the
 * kernel that stages a pack in shared memory (when it fits).
 *
 * b200fem_packed_mesh_create builds such a packing from plain connectivity (what smesh::PackedMesh<T>::create(mesh, {}, true)
 * does inside FunctionSpace::initialize_packed_mesh, frontend/sfem_FunctionSpace.cpp:236-239; smesh is absent from the reference
 * checkout, so the node order inside a pack is ours): nodes are RENUMBERED, `new_to_old[i]` = input id of packed node i -- vectors
 * and points have to be permuted with it. n_elements_per_pack <= 0 picks 1024. Arrays returned by _arrays live until _destroy. */
typedef struct b200fem_packed_mesh_s *b200fem_packed_mesh_t;
int b200fem_packed_mesh_create(b200fem_packed_mesh_t *pm, int nodes_per_element, ptrdiff_t nelements, ptrdiff_t nnodes,
                               const int32_t *const *elements, ptrdiff_t n_elements_per_pack);
int b200fem_packed_mesh_destroy(b200fem_packed_mesh_t pm);
int b200fem_packed_mesh_sizes(b200fem_packed_mesh_t pm, ptrdiff_t *n_packs, ptrdiff_t *n_elements_per_pack,
                              ptrdiff_t *max_nodes_per_pack, ptrdiff_t *n_ghost_entries);
int b200fem_packed_mesh_arrays(b200fem_packed_mesh_t pm, const uint16_t *const **elements, const ptrdiff_t **owned_nodes_ptr,
                               const ptrdiff_t **n_shared, const ptrdiff_t **ghost_ptr, const int32_t **ghost_idx,
                               const int32_t **new_to_old);
/* The reference's atomics-free two-pass apply (SFEM_PACKED_TWO_PASS -> PackedLaplacianApply::apply_two_pass,
 * frontend/ops/sfem_PackedLaplacian.cpp:178-262): pass 1 writes the owned nodes of every pack plainly and parks the ghost results
 * in a buffer, pass 2 adds, per shared node, its buffer entries in a fixed order -- no global atomics. Hand over
 * packed->ghost_reduce_ptr / _idx / _dest (`b200fem_packed_mesh_ghost_reduce` for our packer) with _set_two_pass, then select it
 * with option "packed_two_pass" = 1. Packs must fit shared memory (2 x max_nodes_per_pack scalars <= 200 KB). */
int b200fem_packed_mesh_ghost_reduce(b200fem_packed_mesh_t pm, ptrdiff_t *n_rows, const ptrdiff_t **ghost_reduce_ptr,
                                     const ptrdiff_t **ghost_reduce_idx, const int32_t **ghost_reduce_dest);
int b200fem_packed_laplacian_set_two_pass(b200fem_op_t op, ptrdiff_t n_rows, const ptrdiff_t *ghost_reduce_ptr,
                                          const ptrdiff_t *ghost_reduce_idx, const int32_t *ghost_reduce_dest);
int b200fem_packed_laplacian_create(b200fem_op_t *op, int element_type, ptrdiff_t n_packs, ptrdiff_t n_elements_per_pack,
                                    ptrdiff_t nelements, ptrdiff_t nnodes, ptrdiff_t max_nodes_per_pack,
                                    const uint16_t *const *elements, const float *fff, const ptrdiff_t *owned_nodes_ptr,
                                    const ptrdiff_t *n_shared, const ptrdiff_t *ghost_ptr, const int32_t *ghost_idx, int real_type);
/* VectorLaplacian on hex8: the scalar Laplacian on every component of an AoS field [nnodes * vector_size]
 * (VectorLaplacian::apply -> vector_laplacian_apply_opt -> affine_hex8_vector_laplacian_apply_fff, frontend/ops/
 * sfem_VectorLaplacian.cpp:191-206, operators/vector_laplacian.cpp:41-66, operators/hex8/hex8_vector_laplacian.cpp:78-138);
 * fff as for b200fem_laplacian_create.
 * The semi-structured twin, b200fem_sshex8_vector_laplacian_create <- affine_sshex8_vector_laplacian_apply_fff
 * (operators/sshex8/sshex8_vector_laplacian.cpp:16-97): `elements` / `fff` / `level` as for b200fem_laplacian_create(SSHEX8);
 * the scalar stencil kernel runs component by component on plan-owned work vectors (2 x nnodes scalars). */
int b200fem_vector_laplacian_create(b200fem_op_t *op, int element_type, ptrdiff_t nelements, ptrdiff_t nnodes,
                                    const int32_t *const *elements, const float *fff, int vector_size, int real_type);
int b200fem_sshex8_vector_laplacian_create(b200fem_op_t *op, int level, ptrdiff_t nelements, ptrdiff_t nnodes,
                                           const int32_t *const *elements, const float *fff, int vector_size, int real_type);
/* LumpedMass on sshex8 <- affine_sshex8_mass_lumped (operators/sshex8/sshex8_mass.cpp:11-115; assemble_lumped_mass,
 * operators/mass.cpp:74-77): `diag` only (apply fails: apply_mass has no semi-structured case). `elements` = the (level+1)^3
 * lattice rows, `points[3][nnodes]` read at the 8 macro corners. */
int b200fem_sshex8_lumped_mass_create(b200fem_op_t *op, int level, ptrdiff_t nelements, ptrdiff_t nnodes,
                                      const int32_t *const *elements, const float *const *points, int real_type);
int b200fem_mass_create(b200fem_op_t *op, int element_type, ptrdiff_t nelements, ptrdiff_t nnodes, const int32_t *const *elements,
                        const float *const *points, int real_type);
/* Stored-element-matrix ("EM") semi-structured operators: ss:em:Laplacian / ss:em:LinearElasticity
 * (frontend/ops/sfem_SemiStructuredEMLaplacian.cpp, sfem_SemiStructuredEMLinearElasticity.cpp; kernels
 * operators/stencil/sshex8_stencil_element_matrix_apply.cpp:10-160, GPU twins GPUEMOp / GPUEMWarpOp over
 * operators/sshex8/cuda/cu_sshex8_elemental_matrix.cu:226-1053). `element_matrix` is HOST memory of `real_type`,
 * [nelements][(8 bs)^2] row-major, VTK corner order, components blocked [c0 x 8 | c1 x 8 | c2 x 8]; applied to every micro
 * element of the macro element as out_j += sum_i u_i M[i][j] (operators/sshex8/sshex8.hpp:44-86). block_size 1 or 3.
 * The two builders replace sshex8_laplacian_element_matrix (operators/sshex8/sshex8_laplacian.cpp:595-634; from the macro FFF,
 * scaled by h = 1/level) and sshex8_linear_elasticity_element_matrix (sshex8_linear_elasticity.cpp:1258-1310; from the macro
 * adjugate / determinant); with the first one the EM operator is affine_sshex8_laplacian_apply (sshex8_laplacian.cpp:184-331),
 * the element-matrix formulation of the affine Laplacian. `values_host`: [nelements * 64] / [nelements * 576] of real_type. */
int b200fem_sshex8_em_create(b200fem_op_t *op, int level, ptrdiff_t nelements, ptrdiff_t nnodes, const int32_t *const *elements,
                             int block_size, const void *element_matrix, int real_type);
int b200fem_sshex8_laplacian_element_matrix(int level, ptrdiff_t nelements, const float *fff, int real_type, void *values_host);
int b200fem_sshex8_linear_elasticity_element_matrix(int level, ptrdiff_t nelements, const float *adjugate,
                                                    const float *determinant, double mu, double lambda, int real_type,
                                                    void *values_host);
int b200fem_op_destroy(b200fem_op_t op);

/* y += A x  (sfem::Op::apply(x, h, out), frontend/ops/sfem_Op.hpp:214; Operator<T>::apply,
 * algebra/sfem_Operator.hpp:16) */
int b200fem_op_apply(b200fem_op_t op, const void *x, void *y, void *stream);
/* y = A x : same result as zeros(y) followed by apply, without reading y where no other element
 * contributes. This is the fused form the device CG / Chebyshev below use. */
int b200fem_op_apply_overwrite(b200fem_op_t op, const void *x, void *y, void *stream);
/* The same over the element range [e_begin, e_end) only, for overlapping the halo exchange of a distributed apply
 * with the elements that do not read ghost values ("order interface-adjacent elements last so
 * interior elements can run while the halo is in flight"). mode 0: y += A_range x. mode 1: first part of an
 * overwriting apply (zeroes what b200fem_op_apply_overwrite zeroes, then as mode 2). mode 2: continuation of an
 * overwriting apply. The union of the ranges of one apply must be [0, nelements). Operators that do not implement
 * ranges return B200FEM_FAILURE (the caller then uses the whole-mesh entry points). */
int b200fem_op_apply_range(b200fem_op_t op, const void *x, void *y, ptrdiff_t e_begin, ptrdiff_t e_end, int mode,
                           void *stream);
/* y = A x in ONE launch whose elements >= e_dependent wait, inside the kernel, until *ready_flag >= ready_value (acquire,
 * GPU scope) before they read x: the ghost-dependent elements of a distributed apply, ordered last, start when the
 * halo gather (b200fem_halo_gather_deferred_begin / _launch, on another stream) has filled the ghost block, while
 * everything else runs at once. The launch leaves one SM free for that other stream. Whatever sets the flag must not need a
 * FIRST launch of its kernel while this one waits (lazy module loading may synchronise with running kernels: the wait would
 * run into its 2 s bound; b200fem_halo_gather_deferred_launch goes first for that reason). Operators without the in-kernel wait
 * return B200FEM_FAILURE and enqueue nothing (b200fem_op_get_option "ghost_wait" tells beforehand). */
int b200fem_op_apply_overwrite_after(b200fem_op_t op, const void *x, void *y, ptrdiff_t e_dependent,
                                     const unsigned long long *ready_flag, unsigned long long ready_value, void *stream);
/* diag += diag(A)  (Op::hessian_diag, frontend/ops/sfem_Op.hpp:194; laplacian_diag,
 * operators/laplacian.cpp:156-195; affine_hex8_linear_elasticity_diag,
 * operators/hex8/hex8_linear_elasticity.cpp:1043) */
int b200fem_op_diag(b200fem_op_t op, void *diag, void *stream);
/* 6 entries per node of the symmetric 3x3 block diagonal, AoS [nnodes*6] order
 * (00,01,02,11,12,22)  (Op::hessian_block_diag_sym, frontend/ops/sfem_Op.hpp:138) */
int b200fem_op_block_diag_sym(b200fem_op_t op, void *bdiag, void *stream);

ptrdiff_t b200fem_op_rows(b200fem_op_t op); /* nnodes * block_size */
ptrdiff_t b200fem_op_nelements(b200fem_op_t op); /* what b200fem_op_apply_range's element ranges count over */
int b200fem_op_block_size(b200fem_op_t op);
int b200fem_op_real_type(b200fem_op_t op);
/* Canonical algorithmic bytes per apply: ne*(nxe*4 + geometry bytes) + rows*2*sizeof(real)
 * (the reference's own accounting, notebook/SEMISTRUCTURED.md:20). */
double b200fem_op_algorithmic_bytes(b200fem_op_t op);
/* Bytes the plan actually streams per apply (compressed connectivity, y read when accumulating). */
double b200fem_op_streamed_bytes(b200fem_op_t op, int overwrite);
int b200fem_op_gpu_launches_per_apply(b200fem_op_t op, int overwrite);
/* Tuning switches (all optional; defaults are the measured best). Common: "scatter" (B200FEM_SCATTER_*).
 * hex8 linear elasticity: "patch_elements" (128 [default] | 256 | 512 elements per patch; the plan is rebuilt on the next apply).
 * sshex8 Laplacian: "use_ws" (1 = warp-specialised level-8 kernel where it is the faster one [fp64], 2 = wherever it
 * applies, 0 = never), "ws_cfg" (ring / staging depth variant), "stream_cfg", "use_tma", "explicit_index". */
int b200fem_op_set_option(b200fem_op_t op, const char *key, int value);
int b200fem_op_get_option(b200fem_op_t op, const char *key, int *value);

/* ------------------------------------------------------------------------------------------
 * Dirichlet constraints on device vectors (operators/boundary_conditions/dirichlet.cpp:95-105,
 * 247-258; CUDA twins cuda/cu_boundary_condition.cu:31,79). `idx` is a DEVICE int32 list of node
 * ids; the dof touched is idx[i]*block_size + component.
 * ---------------------------------------------------------------------------------------- */
int b200fem_constraint_nodes_copy_vec(int real_type, ptrdiff_t n, const int32_t *idx, int block_size, int component,
                                      const void *source, void *dest, void *stream);
int b200fem_constraint_nodes_to_value_vec(int real_type, ptrdiff_t n, const int32_t *idx, int block_size,
                                          int component, double value, void *values, void *stream);

/* ------------------------------------------------------------------------------------------
 * Vector kernels = the 13 virtuals of sfem::BLAS<T> (algebra/sfem_tpl_blas.hpp:15-33), semantics
 * of OpenMP_BLAS (algebra/openmp/sfem_openmp_blas.hpp:22-123). Scalars cross the ABI as double.
 * dot / norm2 accumulate in fp64 regardless of real_type.
 * `*_async` variants leave the result in DEVICE memory (one double) and do not synchronise.
 * ---------------------------------------------------------------------------------------- */
int b200fem_blas_zeros(int real_type, ptrdiff_t n, void *x, void *stream);
int b200fem_blas_values(int real_type, ptrdiff_t n, double value, void *x, void *stream);
int b200fem_blas_copy(int real_type, ptrdiff_t n, const void *src, void *dest, void *stream);
int b200fem_blas_scal(int real_type, ptrdiff_t n, double alpha, void *x, void *stream);
int b200fem_blas_axpy(int real_type, ptrdiff_t n, double alpha, const void *x, void *y, void *stream);
int b200fem_blas_axpby(int real_type, ptrdiff_t n, double alpha, const void *x, double beta, void *y, void *stream);
int b200fem_blas_zaxpby(int real_type, ptrdiff_t n, double alpha, const void *x, double beta, const void *y, void *z,
                        void *stream);
int b200fem_blas_xypaz(int real_type, ptrdiff_t n, const void *x, const void *y, double alpha, void *z,
                       void *stream);
int b200fem_blas_reciprocal(int real_type, ptrdiff_t n, double alpha, void *x, void *stream);
int b200fem_blas_dot(int real_type, ptrdiff_t n, const void *x, const void *y, double *result_host, void *stream);
int b200fem_blas_norm2(int real_type, ptrdiff_t n, const void *x, double *result_host, void *stream);
int b200fem_blas_dot_async(int real_type, ptrdiff_t n, const void *x, const void *y, double *result_dev,
                           void *stream);

/* Fused solver updates with the reference's recurrences:
 *   cg_update_x_r_rtr : x += alpha p; r -= alpha Ap; *rtr = (r,r)   (algebra/sfem_cg.hpp:23-44)
 *   cg_update_p       : p = z + beta p                               (algebra/sfem_cg.hpp:64-74)
 *   cheb3_update      : p = beta p - rhs + Ax ... see algebra/sfem_Chebyshev3.hpp:35-52
 * rtr is returned on the host (synchronises the stream). */
int b200fem_cg_update_x_r_rtr(int real_type, ptrdiff_t n, double alpha, const void *p, const void *Ap, void *x,
                              void *r, double *rtr_host, void *stream);
/* The same update with the reduction left in DEVICE memory (one double) and no synchronisation: the distributed CG
 * sums it over the ranks on the device (b200fem_comm_sum_to_host) and reads one scalar per iteration. */
int b200fem_cg_update_x_r_rtr_async(int real_type, ptrdiff_t n, double alpha, const void *p, const void *Ap, void *x,
                                    void *r, double *rtr_dev, void *stream);
int b200fem_cg_update_p(int real_type, ptrdiff_t n, double beta, const void *z, void *p, void *stream);
/* p = beta p - rhs + Ax ; x -= alpha p   (chebyshev3_host_::iteration_update) */
int b200fem_cheb3_update(int real_type, ptrdiff_t n, double alpha, double beta, const void *rhs, const void *Ax,
                         void *p, void *x, void *stream);

/* ------------------------------------------------------------------------------------------
 * Halo exchange helpers. The reference's distributed apply (frontend/sfem_ParallelMatrixFreeOperator.cpp:
 * 78-105) keeps every rank's nodes as [owned | ghosts + aura] and pulls the ghost values from their owners
 * before the local apply (smesh::Exchange::gather, source unavailable). On the GPU the owner side packs the
 * requested owned values into one contiguous send buffer per apply (b200fem_halo_pack), NCCL moves the
 * buffers, and the receiver's ghost block is contiguous so no unpack is needed. unpack_add is the reverse
 * path (smesh::Exchange::scatter_add): dst[idx[i]] += src[i]; the indices of one call must be distinct.
 * `idx` is a DEVICE int32 list of local node ids, values are block_size components per node.
 * ---------------------------------------------------------------------------------------- */
int b200fem_halo_pack(int real_type, ptrdiff_t n, const int32_t *idx, int block_size, const void *src, void *dst,
                      void *stream);
int b200fem_halo_unpack_add(int real_type, ptrdiff_t n, const int32_t *idx, int block_size, const void *src,
                            void *dst, void *stream);

/* ------------------------------------------------------------------------------------------
 * Communicator and halo exchange: the distributed half of the path, one process per GPU.
 * Replaces what the reference takes from smesh (source unavailable): Communicator::sum behind every dot product of
 * ParallelConjugateGradient (algebra/parallel/sfem_parallel_cg.hpp:169-172, 238, 290) and Exchange::gather in
 * ParallelMatrixFreeOperator::apply (frontend/sfem_ParallelMatrixFreeOperator.cpp:78-105).
 *   b200fem_comm_unique_id : rank 0 creates the 128-byte id (ncclGetUniqueId) and the launcher hands it to every rank
 *                            (MPI_Bcast in the reference's world; a torchrun store / broadcast in bench.py).
 *   b200fem_comm_create    : ncclCommInitRank on the CURRENT device; world == 1 needs no id and never touches NCCL.
 *   b200fem_comm_sum_to_host : allreduce(SUM) of n <= 64 DEVICE doubles, then ONE device->host read (synchronises
 *                            `stream`); values_dev is left untouched. This is comm->sum() of the reference with the
 *                            local reduction already on the device.
 *   b200fem_comm_allreduce_sum : the same in place on the device, asynchronous.
 * Halo plan: local nodes are [owned | ghosts], the ghosts received from one peer are a contiguous range
 * [recv_begin, recv_begin + recv_count) (ghosts sorted by owner); send_idx lists the OWNED local ids each peer
 * needs, in the order of that peer's range (CSR over send_peers through send_offsets). gather(x) fills the ghost
 * block of the device vector x ([n_local * block_size], any allocation) from the owners, stream-ordered.
 * Transports: PUSH = the owner's pack kernel stores straight into the peer's receive block over NVLink (cudaIpc-mapped),
 * flags with acquire / release at system scope, every spin bounded by B200FEM_COMM_TIMEOUT_S (default 60 s; a time-out
 * raises b200fem_comm_error and fails the next reduction instead of hanging the GPU); NCCL = pack + grouped
 * ncclSend / ncclRecv into the ghost ranges. AUTO takes PUSH when every rank can map its peers.
 * create / destroy are collective over the communicator.
 * ---------------------------------------------------------------------------------------- */
#define B200FEM_UNIQUE_ID_BYTES 128
#define B200FEM_HALO_AUTO 0
#define B200FEM_HALO_PUSH 1
#define B200FEM_HALO_NCCL 2
#define B200FEM_REDUCE_SUM 0
#define B200FEM_REDUCE_MAX 1
typedef struct b200fem_comm_s *b200fem_comm_t;
typedef struct b200fem_halo_s *b200fem_halo_t;
int b200fem_comm_unique_id(void *id_out /* [B200FEM_UNIQUE_ID_BYTES] */);
int b200fem_comm_create(b200fem_comm_t *comm, int world, int rank, const void *unique_id);
int b200fem_comm_destroy(b200fem_comm_t comm);
int b200fem_comm_rank(b200fem_comm_t comm);
int b200fem_comm_size(b200fem_comm_t comm);
int b200fem_comm_error(b200fem_comm_t comm); /* 1 after a halo wait timed out */
int b200fem_comm_allreduce_sum(b200fem_comm_t comm, double *values_dev, int n, void *stream);
int b200fem_comm_sum_to_host(b200fem_comm_t comm, const double *values_dev, int n, double *values_host, void *stream);
int b200fem_comm_allreduce_host(b200fem_comm_t comm, double *values_host, int n, int op, void *stream);
int b200fem_comm_barrier(b200fem_comm_t comm, void *stream);
int b200fem_halo_create(b200fem_halo_t *halo, b200fem_comm_t comm, int real_type, int block_size, ptrdiff_t n_owned,
                        ptrdiff_t n_local, int n_send_peers, const int *send_peers, const ptrdiff_t *send_offsets,
                        const int32_t *send_idx, int n_recv_peers, const int *recv_peers, const ptrdiff_t *recv_begin,
                        const ptrdiff_t *recv_count, int transport);
/* Indexed receive: the values received from recv_peers[j] are those of the local ghost nodes recv_idx[recv_offsets[j] ..
 * recv_offsets[j+1]) (counts as given to b200fem_halo_create) instead of a contiguous range. For exchanging only the ghosts
 * that can reach an owned row: a matrix-free apply reads a ghost value only through an element that also holds an owned
 * node, and on a semi-structured mesh that is one lattice plane of the aura layer, not all L of them (the reference's
 * Exchange gathers every node of the aura elements). Call on every rank or on none, before the first gather. */
int b200fem_halo_set_recv_index(b200fem_halo_t halo, const ptrdiff_t *recv_offsets, const int32_t *recv_idx);
int b200fem_halo_destroy(b200fem_halo_t halo);
int b200fem_halo_gather(b200fem_halo_t halo, void *x_local, void *stream);
/* The gather in two halves, for a distributed apply that runs its ghost-independent elements in between
 * (b200fem_op_apply_range): begin = the owner side (push over NVLink / pack + grouped NCCL calls),
 * end = the receiver side (wait for the peers' arrival flags, copy into the ghost block; nothing to do for NCCL).
 * begin + end on one stream is b200fem_halo_gather. */
int b200fem_halo_gather_begin(b200fem_halo_t halo, void *x_local, void *stream);
int b200fem_halo_gather_end(b200fem_halo_t halo, void *x_local, void *stream);
/* The whole gather on the halo's OWN high-priority stream, for an apply kernel that waits for the ghost block itself:
 *   deferred_begin  : orders that stream behind what `stream` holds so far and returns a device word *ready_flag that will
 *                     reach *ready_value (release, GPU scope) once the ghost block of x is complete; launches nothing.
 *   deferred_launch : push over NVLink + wait for the peers + copy into the ghost block, on the halo's stream.
 *   b200fem_op_apply_overwrite_after(op, x, y, e_dependent, flag, value, stream) : the apply, one SM left free.
 * (Measured on 2 GPUs at C3: gather launched before the apply 0.784 ms per step, after it on a low-priority stream 1.0 - 1.1 ms
 * -- the gather's blocks then queue behind the apply's persistent CTAs and every peer waits for the late push; gather in front
 * of the apply on one stream: 0.80 - 0.81 ms.)
 * Neither the push nor the wait for the peers then sits in front of the apply. *ready_flag == NULL after deferred_begin: the
 * gather has been enqueued on `stream` in the ordinary way (one rank, NCCL transport, nothing to receive); do not call
 * deferred_launch. The next operation that WRITES x on `stream` must come after the apply (any use of its result does). */
int b200fem_halo_gather_deferred_begin(b200fem_halo_t halo, void *x_local, void *stream, const unsigned long long **ready_flag,
                                       unsigned long long *ready_value);
int b200fem_halo_gather_deferred_launch(b200fem_halo_t halo, void *x_local);
int b200fem_halo_transport(b200fem_halo_t halo); /* the transport in use */
ptrdiff_t b200fem_halo_values_sent(b200fem_halo_t halo);
int b200fem_halo_launches_per_gather(b200fem_halo_t halo);

/* ------------------------------------------------------------------------------------------
 * Multigrid transfer between two lattices of one semi-structured macro mesh.
 * Replaces sshex8_restrict / sshex8_prolongate (operators/sshex8/sshex8_interpolate.cpp:633-775, 259-615)
 * and, with coarse_level = 1, sshex8_hierarchical_restriction / _prolongation (:12-127, 129-237), as wired by
 * create_hierarchical_prolongation / _restriction (frontend/sfem_API.hpp:503-584). Arguments follow the
 * reference: `*_elements` is HOST SoA connectivity with ((level*stride)+1)^3 rows; lattice node (x,y,z) of
 * the level-`level` lattice is row sshex8_lidx(level*stride, x*stride, y*stride, z*stride)
 * (operators/sshex8/sshex8.hpp:18). The plan also holds the element-to-node incidence count of the fine
 * lattice (sshex8_element_node_incidence_count, sshex8_interpolate.cpp:239-257), uint16 per fine node.
 *   restrict  : to[coarse] += W (from[fine] / count[fine])      (caller zeroes `to`, as in the reference)
 *   prolongate: to[fine]    = W^T from[coarse]                   (every fine lattice node is written)
 * Vectors are DEVICE arrays, AoS with `vec_size` components per node.
 * ---------------------------------------------------------------------------------------- */
typedef struct b200fem_transfer_s *b200fem_transfer_t;
int b200fem_sshex8_transfer_create(b200fem_transfer_t *t, ptrdiff_t nelements, int fine_level, int fine_stride,
                                   const int32_t *const *fine_elements, ptrdiff_t n_fine_nodes, int coarse_level,
                                   int coarse_stride, const int32_t *const *coarse_elements, ptrdiff_t n_coarse_nodes);
int b200fem_transfer_destroy(b200fem_transfer_t t);
ptrdiff_t b200fem_transfer_fine_nodes(b200fem_transfer_t t);
ptrdiff_t b200fem_transfer_coarse_nodes(b200fem_transfer_t t);
int b200fem_transfer_incidence_count(b200fem_transfer_t t, uint16_t *count_host /* [n_fine_nodes] */);
/* Distributed meshes: the plan counts over the LOCAL elements, which is the global count for owned nodes only (every
 * element touching an owned node is local); the caller completes the ghost entries from their owners and hands the
 * full table back. */
int b200fem_transfer_set_incidence_count(b200fem_transfer_t t, const uint16_t *count_host /* [n_fine_nodes] */);
int b200fem_transfer_restrict(b200fem_transfer_t t, int real_type, int vec_size, const void *from, void *to,
                              void *stream);
int b200fem_transfer_prolongate(b200fem_transfer_t t, int real_type, int vec_size, const void *from, void *to,
                                void *stream);

/* ------------------------------------------------------------------------------------------
 * Smoother kernels. ShiftableJacobi (algebra/sfem_ShiftableJacobi.hpp:17-57) needs nothing beyond
 * b200fem_blas_copy / _axpy / _reciprocal / _xypaz. ShiftableBlockSymJacobi (:67-144; kernels
 * algebra/openmp/sfem_openmp_ShiftableJacobi.hpp:30-226):
 *   setup: inv_diag[b] = relaxation * inverse( sym3x3(sym_diag6[b] (+ shift6[b])) with the rows of constrained
 *          dofs replaced by identity rows )        -- sym_diag_to_diag, add_sym_diag_to_diag, apply_mask,
 *          inplace_invert, scal in one pass. `shift6` may be NULL. `constraints_mask` is a DEVICE bit mask in
 *          the reference's mask_t layout (bit i%8 of byte i/8 for dof i, algebra/sfem_mask.hpp:13-26) or NULL.
 *   apply: y[b*3+i] += sum_j inv_diag[b*9+i*3+j] x[b*3+j]
 * ---------------------------------------------------------------------------------------- */
int b200fem_block_sym_jacobi_setup(int real_type, ptrdiff_t n_blocks, const void *sym_diag6, const void *shift6,
                                   const void *constraints_mask, double relaxation, void *inv_diag9, void *stream);
int b200fem_block_jacobi_apply(int real_type, ptrdiff_t n_blocks, const void *inv_diag9, const void *x, void *y,
                               void *stream);
/* MixedPrecisionShiftableBlockSymJacobi<double, float> (algebra/sfem_MixedPrecisionShiftableBlockSymJacobi.hpp:13-108): the
 * fp64 block diagonal is converted to fp32, masked, inverted and scaled in fp32; the apply reads fp32 blocks (36 B per node
 * instead of 72 B) with fp64 vectors. */
int b200fem_block_sym_jacobi_setup_mixed(ptrdiff_t n_blocks, const double *sym_diag6, const double *shift6,
                                         const void *constraints_mask, double relaxation, float *inv_diag9, void *stream);
int b200fem_block_jacobi_apply_mixed(ptrdiff_t n_blocks, const float *inv_diag9, const double *x, double *y, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200FEM_H */
