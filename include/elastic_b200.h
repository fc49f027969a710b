/*
 * elastic_b200.h -- C ABI of libelastic_b200.so, the B200-native (sm_100a) replacement for the two
 * data-parallel hot paths of qianyizh/ElasticReconstruction:
 *
 *   TSDF integration   Integrate/TSDFVolume.{h,cpp}, TSDFVolumeUnit.{h,cpp}, ControlGrid.{h,cpp},
 *                      IntegrateApp.cpp:190-269
 *   pairwise ICP       BuildCorrespondence/CorresApp.cpp:112-319 (+ the PCL 1.7 calls it makes)
 *
 * The reference has no plugin/FFI surface: its boundary is two executables.  The entry
 * points below are cut at the granularity of the reference's own methods so that each has an oracle
 * counterpart; the drop-in executables (apps/Integrate, apps/BuildCorrespondence) are thin C++ hosts
 * over exactly these calls.  Each declaration cites the reference code it replaces (paths relative to
 * the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; er_last_error() returns the message of the
 *     last failure on the calling thread.  No exception crosses the ABI.
 *   - plain pointers and sizes only; the caller owns every host buffer.  Pointers named d_* are CUDA
 *     device pointers on the handle's device, everything else is host memory.
 *   - matrices are row-major double[16], camera->world like the reference's .log files
 *     (Integrate/TSDFVolume.cpp:49).
 *   - a handle is bound to one CUDA device and one stream and is not thread-safe; distinct handles are.
 *   - there is no CPU fallback: creating a handle without a usable CUDA device fails.
 */
#ifndef ELASTIC_B200_H_
#define ELASTIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ER_OK 0
#define ER_ERR_CUDA (-1)
#define ER_ERR_ARG (-2)
#define ER_ERR_CAPACITY (-3)

const char* er_last_error(void);
/* Library/ABI version and the CUDA arch the kernels were built for ("sm_100a"). */
int er_version(void);
const char* er_build_arch(void);

/* ------------------------------------------------------------------------------------------------
 * Host-side 4x4 helpers with the reference's (Eigen 3.1.2) evaluation order, exported because the
 * hosts compose trajectories with them:  traj[i*I+j] = pose[i] * seg[i*I+j]  (IntegrateApp.cpp:73).
 * ---------------------------------------------------------------------------------------------- */
void er_mat4_mul(const double a[16], const double b[16], double out[16]);
void er_mat4_inverse(const double m[16], double out[16]);      /* Matrix4d::inverse(), TSDFVolume.cpp:40 */

/* Page-locked host memory for frame buffers handed to the *_batch calls: host->device copies from pageable memory are
 * staged by the driver at a fraction of the link rate.  The pages are placed on the NUMA node of the calling thread's
 * current CUDA device (cudaSetDevice first).  (The reference keeps frames in std::vector, IntegrateApp.h.) */
void* er_host_alloc(size_t bytes);
void er_host_free(void* p);

/* ================================================================================================
 * TSDF volume  (class TSDFVolume, Integrate/TSDFVolume.h:14-69)
 * ============================================================================================== */
typedef struct er_tsdf er_tsdf;

/* TSDFVolume::TSDFVolume (TSDFVolume.cpp:6-13).  voxel/trunc <= 0 select the reference constants
 * 3.0/512 m and 0.03 m; block_res must be 64 (TSDFVolume.cpp:57); max_blocks sizes the device pool of
 * 64^3 (sdf,weight) float2 blocks (2 MiB each; 0 = 1024). */
er_tsdf* er_tsdf_create(int device, int cols, int rows, double voxel, int block_res, double trunc, size_t max_blocks);
void er_tsdf_destroy(er_tsdf* v);

/* Use an existing CUDA stream (cudaStream_t cast to void*) for all work of this handle. */
int er_tsdf_set_stream(er_tsdf* v, void* cuda_stream);
/* CameraParam (TSDFVolumeUnit.h:65-96): fx fy cx cy integration_trunc. */
int er_tsdf_set_camera(er_tsdf* v, float fx, float fy, float cx, float cy, float integration_trunc);
/* ControlGrid::Load for `num` fragments (ControlGrid.cpp:15-34): ctr_xyz = num*(res+1)^3*3 floats. */
int er_tsdf_set_control_grids(er_tsdf* v, const float* ctr_xyz, int num, int res, float length);
/* Multi-GPU block partition: this handle integrates only blocks with (xi+yi+zi) % world == rank. */
int er_tsdf_set_partition(er_tsdf* v, int rank, int world);
/* Count updated voxels / touched blocks in the kernels (roofline bookkeeping; small cost). */
int er_tsdf_set_stats(er_tsdf* v, int enable);
int er_tsdf_get_stats(er_tsdf* v, uint64_t* frames, uint64_t* touched_blocks, uint64_t* updated_voxels);
/* With stats enabled: how many (4^3 sub-brick, frame) pairs took the free-space shortcut (every voxel gets tsdf = 1)
 * and how many the exact per-voxel path; every other pair was skipped as provably unchanged. */
int er_tsdf_get_class_stats(er_tsdf* v, uint64_t* one_pairs, uint64_t* full_pairs, uint64_t* full_idle,
                            uint64_t* full_all_one);   /* last two (optional): exact-path pairs that updated nothing / only with tsdf = 1 */

/* CIntegrateApp::Reproject (IntegrateApp.cpp:228-269) on one frame: depth is rewritten in place.
 * Ti = traj[frame-1], T0 = traj[0], A0 = seg_traj[0], Ai = seg_traj[frame-1], chunk = (frame-1)/interval. */
int er_tsdf_reproject(er_tsdf* v, uint16_t* depth_inout, const double Ti[16], const double T0[16],
                      const double A0[16], const double Ai[16], int chunk);

/* TSDFVolume::ScaleDepth + TSDFVolume::Integrate, the pair of calls at IntegrateApp.cpp:224-225.
 * Asynchronous: returns once the frame is staged and the kernels are enqueued. */
int er_tsdf_integrate(er_tsdf* v, const uint16_t* depth, const double T[16]);
/* Reproject + ScaleDepth + Integrate without the host round trip (IntegrateApp.cpp:217-225). */
int er_tsdf_integrate_ctr(er_tsdf* v, const uint16_t* depth, const double Ti[16], const double T0[16],
                          const double A0[16], const double Ai[16], int chunk);
/* The same for nframes consecutive frames (Ti, Ai = nframes*16 doubles, chunk = nframes ints), batched like
 * er_tsdf_integrate_batch: this is how the 5000-frame ControlGrid configuration is fed. */
int er_tsdf_integrate_ctr_batch(er_tsdf* v, const uint16_t* depth, const double* Ti, const double T0[16],
                                const double A0[16], const double* Ai, const int* chunk, int nframes);
/* The same with the frames already resident in HBM (d_depth is a device pointer), and the device time of the
 * re-projection alone over those frames (reproject_kernel + z-buffer reset/convert, event-bracketed, milliseconds). */
int er_tsdf_integrate_ctr_device(er_tsdf* v, const uint16_t* d_depth, const double* Ti, const double T0[16],
                                 const double A0[16], const double* Ai, const int* chunk, int nframes);
int er_tsdf_profile_reproject_device(er_tsdf* v, const uint16_t* d_depth, const double* Ti, const double T0[16],
                                     const double A0[16], const double* Ai, const int* chunk, int nframes, float* ms);
/* nframes consecutive frames from host memory (depth = nframes*rows*cols, T = nframes*16), in order. */
int er_tsdf_integrate_batch(er_tsdf* v, const uint16_t* depth, const double* T, int nframes);
/* Same with the frames already resident in HBM (d_depth is a device pointer). */
int er_tsdf_integrate_device(er_tsdf* v, const uint16_t* d_depth, const double* T, int nframes);
/* Blocks until every frame handed to the *_device calls so far has been read (ScaleDepth is the only reader of the
 * depth images): the caller may then overwrite those device buffers while the integration itself is still running. */
int er_tsdf_inputs_consumed(er_tsdf* v);
/* Profiling variant of er_tsdf_integrate_device: brackets every kernel launch with CUDA events on the
 * handle's stream and returns the summed device time of the touch (ScaleDepth + block activation),
 * classify (conservative sub-brick culling) and integrate (IntegrateVolumeUnit) kernels over the nframes
 * frames, in milliseconds. */
int er_tsdf_profile_device(er_tsdf* v, const uint16_t* d_depth, const double* T, int nframes, float* touch_ms,
                           float* classify_ms, float* integrate_ms);
/* Frames handed to one fused launch triple by the batch calls: 1..32, default 32.  A voxel is loaded once
 * per batch and updated frame by frame in registers, in order. */
int er_tsdf_set_fusion(er_tsdf* v, int frames_per_launch);
int er_tsdf_sync(er_tsdf* v);

/* Inspection for parity tests: allocated blocks, keyed like TSDFVolume::hash_key (TSDFVolume.h:62-64);
 * sdf/weight are the two float[64^3] arrays of a TSDFVolumeUnit, k fastest (TSDFVolume.cpp:92). */
int er_tsdf_num_blocks(er_tsdf* v, int* n);
int er_tsdf_block_keys(er_tsdf* v, int* keys, int cap);
int er_tsdf_read_block(er_tsdf* v, int key, float* sdf, float* weight);
/* TSDFVolume::SaveWorld predicate (TSDFVolume.cpp:104-132): writes up to cap records x,y,z,intensity
 * (global voxel indices as floats, tsdf) and the total count; order is unspecified as in the reference. */
int er_tsdf_extract(er_tsdf* v, float* xyzi, size_t cap, size_t* n);

/* The same without blocking: er_tsdf_extract_begin enqueues the compaction of the volume as it is after every frame handed
 * in so far (frames handed in later do not affect it) and the copy of `cap` rows to xyzi (page-locked memory, er_host_alloc)
 * on a stream of its own, so that the copy overlaps the next frames' host->device copies; er_tsdf_extract_wait blocks until
 * the rows are home and returns their number (if it exceeds cap only the first cap rows were delivered).  One extraction
 * may be outstanding per handle. */
int er_tsdf_extract_begin(er_tsdf* v, float* xyzi, size_t cap);
int er_tsdf_extract_wait(er_tsdf* v, size_t* n);

/* Raw volume checkpoint (lets a later stage -- meshing, a resumed integration -- start from the device volume
 * without the round trip through world.pcd): every allocated block as the reference's TSDFVolumeUnit (key, float sdf[64^3],
 * float weight[64^3], k fastest; TSDFVolumeUnit.h:98-113).  er_tsdf_load adds the blocks of a checkpoint to a volume with the
 * same voxel size and truncation (a block already present is overwritten); a partitioned volume takes only the blocks it
 * owns, so per-rank checkpoints can be merged by loading them all into one unpartitioned volume. */
int er_tsdf_save(er_tsdf* v, const char* path);
int er_tsdf_load(er_tsdf* v, const char* path);

/* Device-side timing on the handle's stream: record up to 16 events, read elapsed milliseconds. */
int er_tsdf_event_record(er_tsdf* v, int slot);
int er_tsdf_event_elapsed_ms(er_tsdf* v, int slot_a, int slot_b, float* ms);
/* Number of kernel launches issued by this handle so far. */
uint64_t er_tsdf_launch_count(er_tsdf* v);

/* ================================================================================================
 * ICP  (CCorresApp::{Registration,FindCorrespondence}, BuildCorrespondence/CorresApp.cpp:112-319)
 * ============================================================================================== */
typedef struct er_cloud er_cloud;
typedef struct er_icp er_icp;

/* One fragment (pcl::PointCloud<PointXYZRGBNormal> after the NaN-normal filter, CorresApp.cpp:93-98):
 * xyz and normals as n*3 floats.  Builds the uniform-grid search structure that stands in for
 * pcl::KdTreeFLANN::setInputCloud (CorresApp.cpp:148,253) -- bounding box on the host, counting sort by cell on the
 * device; cell_size must be >= every search radius later used against this cloud as a target.  The caller's
 * buffers may be reused on return.  Safe to call from several host threads at once (CorresApp.cpp:82 loads the
 * fragments from an OpenMP loop). */
er_cloud* er_cloud_upload(int device, const float* xyz, const float* normals, size_t n, double cell_size);
void er_cloud_free(er_cloud* c);
size_t er_cloud_size(const er_cloud* c);

/* A built cloud as a self-contained device blob (sorted points, normals, original-order copies, cell table): one rank
 * builds a fragment, the others receive meta + blob (NCCL over NVLink) and call er_cloud_from_blob, which copies the blob
 * into an allocation of its own.  d_blob is a device pointer valid while the cloud lives. */
typedef struct er_cloud_meta {
  uint64_t n, blob_bytes, offset[5];
  double cell;
  float org[3], inv_cell, cell_f;
  int32_t rings, dim[3];
} er_cloud_meta;
int er_cloud_get_meta(const er_cloud* c, er_cloud_meta* meta, const void** d_blob);
er_cloud* er_cloud_from_blob(int device, const er_cloud_meta* meta, const void* d_blob);

/* Gate of CCorresApp::Registration (CorresApp.cpp:250-264): cnt = #{k : NN^2(T*src_k) < dist^2}. */
int er_icp_count_inliers(er_cloud* tgt, er_cloud* src, const double T[16], double dist, int* cnt);

/* The same gate for a list of candidate pairs in one launch and one synchronisation (the reference runs the pair loop
 * from 8 OpenMP threads, CorresApp.cpp:220): pair p tests clouds[src_idx[p]] under T[p] against clouds[tgt_idx[p]];
 * T = npairs*16 doubles, cnt = npairs ints. */
int er_icp_count_inliers_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                               const double* T, double dist, int* cnt);

/* pcl::IterativeClosestPoint + TransformationEstimationPointToPlaneLLS as configured at
 * CorresApp.cpp:295-306 (source = src, target = tgt, guess = Tinit cast to float).  fixed_iters != 0
 * disables the convergence tests so that exactly max_iter iterations run (benchmark mode).
 * Tout = final transformation (float math, widened), iters = iterations run, fitness = mean NN^2 of
 * the last iteration's correspondences. */
int er_icp_align(er_cloud* tgt, er_cloud* src, const double Tinit[16], double max_dist, int max_iter, double eps,
                 int fixed_iters, double Tout[16], int* iters, double* fitness);

/* Many pairs in flight on one device: pair p aligns src[src_idx[p]] onto tgt[tgt_idx[p]].
 * Tinit/Tout = npairs*16 doubles; iters/fitness/converged may be NULL. */
int er_icp_align_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                       const double* Tinit, double max_dist, int max_iter, double eps, int fixed_iters,
                       double* Tout, int* iters, double* fitness);

/* CCorresApp::FindCorrespondence inner loop (CorresApp.cpp:145-161, 186-208): pairs = up to cap
 * (target index, source index) rows in ascending source order, n = their number; info36 (optional) =
 * sum A^T A over the correspondences on the untransformed source points, row-major 6x6. */
int er_icp_correspondences(er_cloud* tgt, er_cloud* src, const double T[16], double dist, double normal_cos,
                           int32_t* pairs, size_t cap, size_t* n, double info36[36]);

/* The same for a list of pairs (the pair loop of FindCorrespondence, CorresApp.cpp:121): the rows of pair p are
 * pairs[2*row_start[p] .. 2*row_start[p+1]), row_start has npairs+1 entries; info36 (optional) = npairs*36 doubles.
 * Returns ER_ERR_CAPACITY (with row_start filled in) when more than `cap` rows were found. */
int er_icp_correspondences_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                                 const double* T, double dist, double normal_cos, int32_t* pairs, size_t cap,
                                 size_t* row_start, double* info36);

/* The batched calls keep their device and pinned-host working memory between calls (per device); this frees it. */
int er_icp_release(int device);

/* Timing / accounting for the ICP path (events are recorded on the library's ICP stream of `device`). */
int er_icp_event_record(int device, int slot);
int er_icp_event_elapsed_ms(int device, int slot_a, int slot_b, float* ms);
uint64_t er_icp_launch_count(int device);
int er_icp_sync(int device);

/* ================================================================================================
 * FragmentOptimizer  (class COptApp + class PointCloud, FragmentOptimizer/OptApp.{h,cpp}, PointCloud.{h,cpp};
 * SURVEY.md section 8f #2).  The data-parallel parts -- the per-point lattice coordinates and their updates, and the
 * assembly of the normal equations over all correspondences of all fragment pairs -- run on the device; the control
 * flow of OptimizeSLAC / OptimizeRigid / OptimizeNonrigid, the lattice regulariser (729 small SVDs) and the file
 * formats live in apps/FragmentOptimizer.cpp, a C++ host over exactly these calls.
 *
 * Index conventions are the reference's: fragments 0..num-1 are the RELATIVE indices after blacklisting
 * (OptApp.cpp:31-47); a lattice has (resolution+1)^3 nodes, node (i,j,k) at i + j(res+1) + k(res+1)^2, three doubles
 * each (nper = 3(res+1)^3); the SLAC unknown vector is [6 per fragment | nper] (matrix_size_ = 6 num + nper,
 * OptApp.cpp:428), the rigid one [6 per fragment], the nonrigid one [nper per fragment].
 * ============================================================================================== */
typedef struct er_fragopt er_fragopt;

er_fragopt* er_fragopt_create(int device, int num, int resolution, double length);   /* COptApp members, OptApp.cpp:126-128 */
void er_fragopt_destroy(er_fragopt* h);

/* PointCloud::LoadFromPCDFile / LoadFromXYZNFile after parsing (PointCloud.cpp:21-63): xyzn = n rows of
 * x y z nx ny nz (rows with a NaN normal already dropped by the caller).  GetCoordinate (PointCloud.h:96-177) runs for
 * every row; like the reference, loading stops at the first point outside [0, length)^3 ("Point out of bound"):
 * *n_kept = number of points the cloud then holds (the failing one included, as the zero-filled record the reference
 * leaves behind), and the call still succeeds.  All clouds must be set before er_fragopt_set_pairs. */
int er_fragopt_set_cloud(er_fragopt* h, int l, const float* xyzn, size_t n, size_t* n_kept);

/* COptApp::InitCorrespondences (OptApp.cpp:100-118) after parsing: pair p joins fragments idx0[p] and idx1[p]; its rows
 * (index into cloud idx0[p], index into cloud idx1[p]) are rows[2*row_start[p] .. 2*row_start[p+1]).  The library keeps
 * them on the device ordered by (pair, lattice cell of the first point, lattice cell of the second), the order in which
 * the assembly kernels walk them. */
int er_fragopt_set_pairs(er_fragopt* h, int npairs, const int* idx0, const int* idx1, const size_t* row_start,
                         const int32_t* rows);

/* PointCloud::UpdatePose for every fragment (PointCloud.h:70-83; OptApp.cpp:297,402,443): T = num row-major 4x4
 * doubles, cast to float like pose.cast<float>(), applied to positions and normals. */
int er_fragopt_update_pose(er_fragopt* h, const double* T);
/* PointCloud::UpdateAllNormal (PointCloud.h:37-41; OptApp.cpp:151-153): ctr = num*nper doubles. */
int er_fragopt_update_normals(er_fragopt* h, const double* ctr);
/* PointCloud::UpdateAllPointPN (PointCloud.h:43-51; OptApp.cpp:660-663): expand_ctr = num*nper doubles. */
int er_fragopt_update_points_pn(er_fragopt* h, const double* expand_ctr);
/* Current position and normal of every point of fragment l (n rows of x y z nx ny nz); UpdatePoint under `ctr`
 * (PointCloud.h:85-94, used by SavePoints, OptApp.cpp:905-907) when ctr != NULL: then xyz = the warped position. */
int er_fragopt_get_points(er_fragopt* h, int l, float* xyzn, size_t cap_rows, const double* ctr);

/* Data term of one SLAC iteration (the pair loop of OptimizeSLAC, OptApp.cpp:473-570): rot_t = pose_rot_t_, num
 * row-major 3x3 doubles.  The dense upper triangle of JJ (matrix_size_^2 doubles) stays on the device; Jb
 * (matrix_size_ doubles) and the data error score come back.  One kernel over all correspondences: a 128-thread CTA
 * stages 128 Jacobian rows (60 entries each) in shared memory and accumulates their outer products in 4x4 register
 * tiles; sums leave a CTA only when the 60 column indices change (the sorted order makes that rare). */
int er_fragopt_slac_assemble(er_fragopt* h, const double* rot_t, double* Jb, double* score);
/* Solves (base + JJ_data) x = rhs by a dense Cholesky factorisation on the device (the reference:
 * CholmodSupernodalLLT on thisJJ.sparseView(), OptApp.cpp:574-577,639).  base = the regulariser part, dense row-major
 * matrix_size_^2 doubles of which the upper triangle is read (baseJJ * default_weight plus the six ones of
 * OptApp.cpp:456-464); NULL = the matrix given last time. */
int er_fragopt_slac_solve(er_fragopt* h, const double* base, const double* rhs, double* x);
/* Several GPUs: every device holds a handle with ALL clouds and a share of the pairs (the pair loop of OptimizeSLAC shards like
 * the ICP one); after each handle's er_fragopt_slac_assemble, dst's matrix += src's -- a peer copy of the matrix_size_^2
 * doubles (49 MB for 50 fragments) into a buffer on dst's device and one add kernel there.  Jb and the score are host
 * values the caller adds.  This is the one exchange step of the path. */
int er_fragopt_slac_accumulate(er_fragopt* dst, er_fragopt* src);
/* The assembled data term, upper triangle of a row-major matrix_size_^2 matrix (tests, diagnostics). */
int er_fragopt_slac_get_jj(er_fragopt* h, double* jj);

/* Data term of one rigid iteration (OptApp.cpp:312-385): JJ = upper triangle of a row-major (6 num)^2 matrix,
 * Jb = 6 num doubles, score = sum b^2.  (The six ones every pair adds to the first six diagonal entries,
 * OptApp.cpp:323-325, are the caller's.) */
int er_fragopt_rigid_assemble(er_fragopt* h, double* JJ, double* Jb, double* score);

/* Nonrigid mode (COptApp::OptimizeNonrigid, OptApp.cpp:120-278; unknowns = num * nper lattice coordinates).  The system
 * is kept DENSE on the device (row-major N x N doubles, N = num * nper, upper triangle; 344 MB for 3 fragments, 3.8 GB
 * for 10, 96 GB for 50 -- ER_ERR_CAPACITY when it does not fit) and factorised by cuSOLVER where the reference calls
 * CholmodSupernodalLLT on a sparse matrix.
 *   assemble  the pair loop (OptApp.cpp:159-206): sum of row^T row over all correspondences, rows of 48 entries
 *             (pi.val * weight * n_i on fragment i's lattice, -pj.val * weight * n_i on fragment j's); call
 *             er_fragopt_update_normals first (OptApp.cpp:151-153)
 *   factor    thisAA = data + baseAA (given as upper-triangle triplets), then the Cholesky factorisation (OptApp.cpp:155,209-211)
 *   solve     ctr = solver.solve(Ab) (OptApp.cpp:262), any number of times per factorisation */
int er_fragopt_nonrigid_assemble(er_fragopt* h, double weight);
int er_fragopt_nonrigid_factor(er_fragopt* h, size_t nnz, const int32_t* rows, const int32_t* cols, const double* vals);
int er_fragopt_nonrigid_solve(er_fragopt* h, const double* rhs, double* x);
/* The assembled data term (N*N doubles, upper triangle), valid between assemble and factor (tests). */
int er_fragopt_nonrigid_get_aa(er_fragopt* h, double* aa);

/* Timing / accounting like the other paths (events on the handle's stream). */
int er_fragopt_event_record(er_fragopt* h, int slot);
int er_fragopt_event_elapsed_ms(er_fragopt* h, int slot_a, int slot_b, float* ms);
uint64_t er_fragopt_launch_count(const er_fragopt* h);
int er_fragopt_sync(er_fragopt* h);
/* Device time of the last er_fragopt_slac_assemble's assembly kernel alone, and the number of segment flushes it made. */
int er_fragopt_last_assemble_stats(er_fragopt* h, float* kernel_ms, uint64_t* rows, uint64_t* flushes);

#ifdef __cplusplus
}
#endif
#endif /* ELASTIC_B200_H_ */
