/*
 * b200_bundle_adjustment.h — C-ABI of the B200-native bundle-adjustment LM solver.
 *
 * Drop-in boundary: a third `BundleAdjustmentBackend` next to CERES and CASPAR
 * (reference: src/colmap/estimators/bundle_adjustment.h:60,212-234 and the factory
 * CreateDefaultBundleAdjuster, bundle_adjustment.cc:314-336).  The C++ adapter
 *  flattens `Reconstruction` into b200ba_problem exactly as
 * CasparBundleAdjuster::BuildFactors does (bundle_adjustment_caspar.cc:104-371) and writes the
 * in/out arrays back (ibid. :767-801).
 *
 * What the solver does (reference behaviour = DefaultBundleAdjuster + ceres::Solve,
 * bundle_adjustment_ceres.cc:574-683): Levenberg-Marquardt with Ceres' trust-region rules, analytic
 * reprojection Jacobians (cost_functions/reprojection_error.h:62-210), quaternion (x) R^3 manifold,
 * Schur elimination of the point blocks and either an exact reduced solve (DENSE_/SPARSE_SCHUR) or
 * PCG with the SCHUR_JACOBI preconditioner (ITERATIVE_SCHUR), all on the GPU in fp64.
 *
 * Conventions: plain pointers and sizes, 0 == success, negative == error, no exceptions.
 * `poses`, `camera_params`, `points` are updated IN PLACE; constant blocks are left bit-identical
 * (bundle_adjustment_test.cc:409-411 relies on this).
 */
#ifndef B200_BUNDLE_ADJUSTMENT_H_
#define B200_BUNDLE_ADJUSTMENT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CameraModelId values (src/colmap/sensor/models.h:90-109): all eighteen models are on the solve path.  The radial
 * pinhole family and its two equidistant-fisheye counterparts (<= 5 parameters) go through hand-written Jacobians and
 * the narrow kernels; the other twelve through forward-mode dual numbers (colmap_b200/csrc/ba_models.cuh) and the
 * wide kernel instantiations. */
enum { B200BA_SIMPLE_PINHOLE = 0, B200BA_PINHOLE = 1, B200BA_SIMPLE_RADIAL = 2, B200BA_RADIAL = 3, B200BA_OPENCV = 4,
       B200BA_OPENCV_FISHEYE = 5, B200BA_FULL_OPENCV = 6, B200BA_FOV = 7, B200BA_SIMPLE_RADIAL_FISHEYE = 8,
       B200BA_RADIAL_FISHEYE = 9, B200BA_THIN_PRISM_FISHEYE = 10, B200BA_RAD_TAN_THIN_PRISM_FISHEYE = 11,
       B200BA_SIMPLE_DIVISION = 12, B200BA_DIVISION = 13, B200BA_SIMPLE_FISHEYE = 14, B200BA_FISHEYE = 15, B200BA_EUCM = 16,
       B200BA_EQUIRECTANGULAR = 17 };

/* ceres::LinearSolverType subset COLMAP selects (bundle_adjustment_ceres.cc:202-212). */
enum { B200BA_AUTO = 0, B200BA_DENSE_SCHUR = 1, B200BA_SPARSE_SCHUR = 2, B200BA_ITERATIVE_SCHUR = 3 };

/* CeresBundleAdjustmentOptions::LossFunctionType (bundle_adjustment_ceres.h:41-43). */
enum { B200BA_LOSS_TRIVIAL = 0, B200BA_LOSS_SOFT_L1 = 1, B200BA_LOSS_CAUCHY = 2, B200BA_LOSS_HUBER = 3 };

/* BundleAdjustmentTerminationType (bundle_adjustment.h:50-57). */
enum { B200BA_CONVERGENCE = 0, B200BA_NO_CONVERGENCE = 1, B200BA_FAILURE = 2 };

/* BundleAdjustmentOptions (bundle_adjustment.h:175-208) + the Ceres solver options COLMAP sets
 * (CeresBundleAdjustmentOptions ctor, bundle_adjustment_ceres.cc:102-116); Ceres defaults otherwise. */
typedef struct b200ba_options {
  int refine_focal_length;              /* 1 */
  int refine_principal_point;           /* 0 */
  int refine_extra_params;              /* 1 */
  int refine_rig_from_world;            /* 1 */
  int refine_points3D;                  /* 1 */
  int constant_rig_from_world_rotation; /* 0 */
  int loss_function_type;               /* TRIVIAL */
  double loss_function_scale;           /* 1.0 */
  int linear_solver_type;               /* AUTO: <=50 images DENSE_SCHUR, <=1000 SPARSE_SCHUR, else ITERATIVE_SCHUR */
  int max_num_iterations;               /* 100 */
  int max_linear_solver_iterations;     /* 200 */
  double function_tolerance;            /* 0 */
  double gradient_tolerance;            /* 1e-4 */
  double parameter_tolerance;           /* 0 */
  double initial_trust_region_radius;   /* 1e4  (Ceres default) */
  double max_trust_region_radius;       /* 1e16 */
  double min_trust_region_radius;       /* 1e-32 */
  double min_relative_decrease;         /* 1e-3 */
  double min_lm_diagonal;               /* 1e-6 */
  double max_lm_diagonal;               /* 1e32 */
  double eta;                           /* 0.1: CG forcing term (q-tolerance) for ITERATIVE_SCHUR */
  int jacobi_scaling;                   /* 1 */
  int gpu_index;                        /* -1 = current device */
  int refine_sensor_from_rig;           /* 1 (BundleAdjustmentOptions::refine_sensor_from_rig, bundle_adjustment.h:187) */
} b200ba_options;

/* Flat problem.  Trivial frames: one pose per image.  Rigs (non-trivial frames, reprojection_error.h:344-420,
 * bundle_adjustment_ceres.cc:753-827): obs_pose_idx indexes the rig_from_world pose of the FRAME (shared by the frame's
 * images) and a camera that is a non-reference sensor carries a sensor_from_rig pose;
 * cam_from_world = sensor_from_rig * rig_from_world. */
typedef struct b200ba_problem {
  int num_poses;
  double* poses;                           /* [7*num_poses] qx qy qz qw tx ty tz (geometry/rigid3.h:46-49), in/out */
  const uint8_t* pose_constant;            /* [num_poses] or NULL */
  const int8_t* pose_fixed_translation_dim;/* [num_poses] -1 | 0..2 (two-cams gauge, bundle_adjustment_ceres.cc:400-416) or NULL */
  int num_cameras;
  const int32_t* camera_model_id;          /* [num_cameras] */
  const int32_t* camera_param_offset;      /* [num_cameras] into camera_params */
  double* camera_params;                   /* in/out */
  const uint8_t* camera_constant;          /* [num_cameras] HasConstantCamIntrinsics, or NULL */
  int64_t num_points;
  double* points;                          /* [3*num_points], in/out */
  const uint8_t* point_constant;           /* [num_points] or NULL */
  int64_t num_observations;
  const int32_t* obs_pose_idx;
  const int32_t* obs_camera_idx;
  const int32_t* obs_point_idx;
  const double* obs_xy;                    /* [2*num_observations] */
  int32_t num_config_images;               /* config.NumImages(): drives the AUTO linear-solver choice (bundle_adjustment_ceres.cc:131,
                                            * 204-210); 0 = the number of poses that appear in observations */
  int32_t num_sensors;                     /* non-reference rig sensors (0 for trivial frames) */
  double* sensor_from_rig;                 /* [7*num_sensors] qx qy qz qw tx ty tz, in/out */
  const uint8_t* sensor_constant;          /* [num_sensors] HasConstantSensorFromRigPose / reference sensor absent (:526-539), or NULL */
  const int32_t* camera_sensor_idx;        /* [num_cameras] index into sensor_from_rig, -1 = reference sensor; NULL = all -1 */
} b200ba_problem;

/* BundleAdjustmentSummary (bundle_adjustment.h:63-74) + the ceres::Solver::Summary fields
 * PrintSolverSummary reports (bundle_adjustment_ceres.cc:1102-1137). */
typedef struct b200ba_summary {
  int termination_type;
  int num_residuals;            /* 2 x observations connected to >= 1 variable block */
  int num_effective_parameters;
  int num_successful_steps;
  int num_unsuccessful_steps;
  int num_linear_solver_iterations; /* total PCG iterations */
  int linear_solver_type_used;
  double initial_cost, final_cost;  /* 1/2 sum rho(r^2) */
  double solve_ms;              /* device time of the LM loop */
  double setup_ms;              /* host flattening + H2D */
  double spmv_ms_total;         /* device time inside the implicit-Schur SpMV kernels */
  int spmv_launches;
  int kernel_launches;
} b200ba_summary;

void b200ba_options_init(b200ba_options* o);

/* FixGaugeWithTwoCamsFromWorld (bundle_adjustment_ceres.cc:308-417) on the flat problem: writes
 * pose_constant_out[num_poses] / fixed_dim_out[num_poses]; pose order = ascending image id.
 * Returns 0, or 1 if no valid pair exists (the reference then falls back to three fixed points). */
int b200ba_fix_gauge_two_cams_from_world(const b200ba_problem* p, const b200ba_options* o, uint8_t* pose_constant_out,
                                         int8_t* fixed_dim_out);

/* BundleAdjuster::Solve. */
int b200ba_solve(const b200ba_options* o, b200ba_problem* p, b200ba_summary* out);

/* Multi-GPU (one process per GPU): points are sharded over the ranks, every rank passes the same poses / cameras and
 * its own points with ALL their observations (so the point-block elimination stays local).  The communicator wraps
 * NCCL: rank 0 creates a 128-byte id, the launcher broadcasts it (torch.distributed / MPI), every rank calls
 * b200ba_comm_init on its CUDA device.  Collectives: one all-reduce of the camera-side vector per PCG iteration plus
 * a few small ones per LM iteration.  poses / camera_params are returned identical on every rank. */
typedef struct b200ba_comm* b200ba_comm_t;
int b200ba_comm_unique_id(void* id128);
int b200ba_comm_init(const void* id128, int rank, int world_size, b200ba_comm_t* out);
/* 1 once the communicator's ranks have mapped each other's symmetric buffers (cudaIpc over NVLink) and the small
 * collectives of a sharded solve run as the library's own one-shot peer-memory all-reduce kernel; 0 = NCCL only (before
 * the first sharded solve, world > 8, no peer access, or B200BA_NO_P2P set). */
int b200ba_comm_peer_memory(b200ba_comm_t comm);
void b200ba_comm_destroy(b200ba_comm_t comm);
int b200ba_solve_sharded(const b200ba_options* o, b200ba_problem* local_shard, b200ba_comm_t comm, b200ba_summary* out);

/* ---- problem assembly from a Reconstruction + BundleAdjustmentConfig ----
 * DefaultBundleAdjuster's constructor (bundle_adjustment_ceres.cc:606-664: AddImageToProblem :688-751, AddPointToProblem
 * :829-888, ParameterizeCameras / ParameterizeRigsAndFrames / ParameterizePoints :419-565, FixGauge :270-417), on flat views of the
 * reference's containers.  The adapter fills the views with plain loops over Reconstruction (no hash-map walks remain on
 * the solve path), calls b200ba_assemble, b200ba_solve on the assembled problem, and copies the three in/out arrays back. */
typedef struct b200ba_scene {
  int num_images;                   /* every image of the reconstruction, ascending image id */
  const uint32_t* image_id;
  const int32_t* image_camera;      /* [num_images] index into the camera arrays */
  const double* cam_from_world;     /* [7*num_images] qx qy qz qw tx ty tz */
  const int64_t* point2D_offset;    /* [num_images+1] into the two arrays below */
  const double* point2D_xy;         /* [2*total] */
  const int64_t* point2D_point3D;   /* [total] index into the point arrays, -1 = no 3D point */
  int num_cameras;
  const int32_t* camera_model_id;
  const int32_t* camera_param_offset;
  const double* camera_params;
  int64_t num_points3D;             /* ascending point3D id */
  const double* xyz;                /* [3*num_points3D] */
  const int64_t* track_offset;      /* [num_points3D+1] */
  const int32_t* track_image;       /* image INDEX of every track element */
  const int32_t* track_point2D;     /* index of the element inside that image's point2D list */
  /* ---- rigs and frames (scene/rig.h, scene/frame.h); num_frames == 0: trivial frames, cam_from_world above is the pose.
   * With frames the pose of an image is sensor_from_rig(camera) * rig_from_world(frame) and cam_from_world is ignored. */
  int num_frames;                   /* ascending frame id */
  const int32_t* image_frame;       /* [num_images] frame index */
  const double* rig_from_world;     /* [7*num_frames] */
  const int32_t* frame_rig;         /* [num_frames] rig index */
  int num_rigs;
  const int32_t* rig_ref_camera;    /* [num_rigs] camera index of the rig's reference sensor */
  const int32_t* camera_rig;        /* [num_cameras] rig the camera is a sensor of */
  const double* camera_sensor_from_rig;   /* [7*num_cameras]; ignored for reference sensors */
} b200ba_scene;

typedef struct b200ba_config {      /* BundleAdjustmentConfig (bundle_adjustment.h:77-151) as flags over the scene arrays */
  const uint8_t* image_in_config;       /* [num_images] */
  const uint8_t* image_constant_pose;   /* HasConstantRigFromWorldPose */
  const uint8_t* camera_constant;       /* [num_cameras] HasConstantCamIntrinsics */
  const uint8_t* point_variable;        /* [num_points3D] AddVariablePoint */
  const uint8_t* point_constant;        /* AddConstantPoint */
  const uint8_t* point_ignored;         /* IgnorePoint */
  int fixed_gauge;                      /* 0 UNSPECIFIED, 1 TWO_CAMS_FROM_WORLD (:308-417), 2 THREE_POINTS (:270-306; points in
                                         * ascending id, the reference walks a hash map) - enum at bundle_adjustment.h:44-48 */
  int min_track_length;                 /* BundleAdjustmentOptions::min_track_length */
  /* with frames (scene.num_frames > 0): image_constant_pose is ignored, these two are read instead */
  const uint8_t* frame_constant_pose;                /* [num_frames] HasConstantRigFromWorldPose */
  const uint8_t* camera_constant_sensor_from_rig;    /* [num_cameras] HasConstantSensorFromRigPose */
} b200ba_config;

typedef struct b200ba_assembly* b200ba_assembly_t;
int b200ba_assemble(const b200ba_options* o, const b200ba_scene* scene, const b200ba_config* config, b200ba_assembly_t* out);
/* the assembled problem; flat pose / camera / point indices equal the scene's image / camera / point indices.  The
 * arrays belong to the assembly; poses, camera_params and points are copies that b200ba_solve updates in place.
 * With frames: pose k = frame k (rig_from_world) for k < num_frames, followed by one constant pose per image outside the
 * config that a config point brought in (a copy of its frame's pose; the reference bakes those observations with a
 * constant pose, bundle_adjustment_ceres.cc:846-878); sensor s = the s-th non-reference camera in ascending camera
 * index (problem->sensor_from_rig, updated in place by the solve). */
b200ba_problem* b200ba_assembly_problem(b200ba_assembly_t a);
void b200ba_assembly_free(b200ba_assembly_t a);

const char* b200ba_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_BUNDLE_ADJUSTMENT_H_ */
