/* hunter_b200.h -- C ABI of libhunter_b200.so: batched NMPC iteration + WeightedWbc QP for the Hunter biped on B200.
 *
 * Every entry point replaces one operator of the reference's per-control-step path; the reference-side
 * bindings (C++ adapters deriving from ocs2::MPC_BASE / legged::WbcBase) are shown in INTEGRATION.md.
 *
 *   hb_wbc_solve_batch      <-> legged::WeightedWbc::update            legged_wbc/src/WeightedWbc.cpp:18-66,
 *                               legged::WbcBase::update               legged_wbc/include/legged_wbc/WbcBase.h:43-44
 *   hb_wbc_qp_batch         <-> qpOASES::QProblem::init + getPrimalSolution   legged_wbc/src/WeightedWbc.cpp:44-55
 *   hb_hierarchical_wbc_solve_batch <-> legged::HierarchicalWbc::update   legged_wbc/src/HierarchicalWbc.cpp:18-31
 *   hb_hoqp_solve_batch     <-> legged::HoQp (null-space cascade)          legged_wbc/src/HoQp.cpp:21-198
 *   hb_mpc_solve_batch      <-> ocs2::MPC_MRT_Interface::advanceMpc -> SqpMpc/SqpSolver::run (one SQP iteration)
 *                               legged_controllers/src/LeggedController.cpp:378-379,406
 *   hb_mpc_cold_start_batch <-> LeggedRobotInitializer::compute       legged_interface/src/initialization/LeggedRobotInitializer.cpp:67-77
 *   hb_policy_eval_batch    <-> MPC_MRT_Interface::evaluatePolicy     legged_controllers/src/LeggedController.cpp:154-156
 *   hb_control_step_batch   <-> LeggedController::update MPC->policy->WBC->torque law   LeggedController.cpp:137-257
 *   hb_resident_cycle_batch <-> SqpSolver::run with its resident primalSolution_ (warm start) + the rest of LeggedController::update
 *   hb_resident_plan_cycle_batch <-> ReferenceManager::preSolverRun (planner on the device) + hb_resident_cycle_batch
 *   hb_estimator_update_batch <-> KalmanFilterEstimate::update        legged_estimation/src/LinearKalmanFilter.cpp:72-185
 *   hb_contact_force_estimate_batch <-> StateEstimateBase::estContactForce   legged_estimation/src/StateEstimateBase.cpp:130-206
 *   hb_joint_command_batch  <-> joint command / torque law            LeggedController.cpp:186-257
 *   hb_plan_references      <-> GaitSchedule tiling + SwingTrajectoryPlanner::update + cmdVelToTargetTrajectories + calculateJointRef
 *   hb_gait_select          <-> SwitchedModelReferenceManager::calculateVelAbs + walkGait/trotGait   :185-249
 *   hb_rbd_to_centroidal_batch <-> CentroidalModelRbdConversions::computeCentroidalStateFromRbdModel  LeggedController.cpp:336
 *   hb_reference_expand_batch  <-> SwitchedModelReferenceManager::modifyReferences (gait tiling, swing planner, target
 *                               interpolation, evaluated on the node grid)   legged_interface/src/SwitchedModelReferenceManager.cpp:136-171
 *
 * Conventions: plain pointers and sizes only; no exceptions cross the ABI. Return 0 on success, <0 on misuse / CUDA error
 * (hb_strerror). Per-instance status words: 0 converged, 1 iteration cap, 2 infeasible / ill-posed, 3 NaN.
 * All arithmetic is IEEE float64 (the reference is all-double: ocs2::scalar_t).
 * Layouts (row-major, instance-major): state x[22] = [h_lin/m, h_ang/m, p, zyx, q_j]; input u[22] = [F0..F3, qj_dot];
 * rbd[32] = [zyx, p, q_j, omega_world, v, qj_dot]; WBC solution sol[38] = [qdd(16), F(12), tau(10)].
 * The *_host entry points take host pointers (pinned or pageable) and copy through the context's staging buffers;
 * the *_dev entry points take device pointers and run asynchronously on the context's stream (hb_sync to wait).
 */
#ifndef HUNTER_B200_H
#define HUNTER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_ctx hb_ctx;

typedef struct {
  int32_t horizon_N;     /* shooting intervals (BASELINE: 100; shipped task.info: ~54)            */
  double dt;             /* node spacing [s] (BASELINE: 0.01; task.info:82 0.015)                 */
  int32_t max_batch;     /* capacity of the context's scratch buffers                             */
  double wbc_rho;        /* Tikhonov weight that defines the least-norm WBC optimum (qpOASES setToMPC regularisation) */
  int32_t qp_max_iter;   /* interior-point iteration cap                                          */
  int32_t line_search_max_trials; /* alpha = 1, 1/2, ... >= 1e-4 -> 14                            */
  double time_horizon;   /* mpc.timeHorizon [s] (task.info:144 0.8); 0 = horizon_N * dt. Used by the event-node grid */
  int32_t event_nodes;   /* 1: the resident cycle discretises [t0, t0 + time_horizon] like ocs2::timeDiscretizationWithEvents
                            (steps of dt, a node on every mode switch, grid re-anchored there, last node = final time); horizon_N is
                            then the node capacity. 0: uniform grid t0 + k dt (BASELINE configs)                          */
  int32_t e2e_chunks;    /* host-pointer cycle calls split the batch into this many chunks pipelined over two streams (copies of one
                            chunk overlap kernels of the other); 0 = automatic                                            */
} hb_config;

typedef struct {
  double alpha;          /* accepted step length (0 = step rejected, iterate kept)                */
  double merit0, merit1; /* merit before / after                                                  */
  double viol0, viol1;   /* sqrt(dynamics SSE + equality SSE) before / after                      */
  double armijo;         /* descent metric of the projected subproblem                            */
  int32_t status;        /* 0 ok, 3 NaN                                                           */
  int32_t n_trials;      /* line-search trials evaluated                                          */
} hb_solve_info;

/* compact per-instance reference description consumed by hb_reference_expand_batch */
#define HB_MAX_EVENTS 32
#define HB_MAX_TARGETS 16
#define HB_MAX_SEGMENTS 24
#define HB_MAX_HORIZON 512   /* hb_create rejects longer horizons (shared-memory staging of one instance's trajectories) */
typedef struct {
  int32_t n_events;                       /* mode schedule: modes[i] holds on (event_times[i-1], event_times[i]] */
  double event_times[HB_MAX_EVENTS];
  int32_t modes[HB_MAX_EVENTS + 1];
  int32_t n_targets;                      /* target trajectory samples (time, state)                              */
  double target_times[HB_MAX_TARGETS];
  double target_states[HB_MAX_TARGETS][22];
  int32_t n_segments[4][3];               /* per contact, per axis: cubic Hermite segments [t0,t1,p0,v0,p1,v1]    */
  double segments[4][3][HB_MAX_SEGMENTS][6];
} hb_reference;

/* inputs of the host-side reference planner (one MPC solve of one instance) */
typedef struct {
  double t0;              /* solver init time                                                             */
  double horizon;         /* final time = t0 + horizon                                                    */
  double time_to_target;  /* TIME_TO_TARGET of the cmd_vel target (TargetTrajectoriesPublisher.cpp:107)   */
  double gait_start;      /* time at which the gait template starts (STANCE before)                       */
  double prev_event;      /* an event time inside the initial stance (< gait_start)                       */
  double x0[22];          /* current observation state                                                    */
  double cmd_vel[4];      /* filtered command: vx, vy, vz, yaw rate (body frame)                          */
  double feet_pos[12];    /* current contact positions in world (hb_contact_positions_batch)              */
  int32_t gait;           /* 0 stance, 1 trot, 2 standing_trot, 3 flying_trot (reference.info:54-118)     */
  int32_t joint_ik;       /* 1: resample the target every 0.15 s and fill joint references by IK (calculateJointRef,
                             SwitchedModelReferenceManager.cpp:251-300); 0: keep the two-sample target with default joints */
} hb_plan_input;

/* state of the speed-based gait selection of one instance (SwitchedModelReferenceManager velAbsHistory_/velAvg_/gaitLevel_);
 * zero-initialise, then set gait_level = -1 ("no template chosen yet") or the level in force */
typedef struct {
  double history[50];
  double vel_avg;
  int32_t head, count;
  int32_t gait_level;
  int32_t reserved;
} hb_gait_selector;

/* joint PD gains of the command law (dynamic_reconfigure parameters, legged_controllers/cfg/Tutorials.cfg:6-16,
 * LeggedController.cpp:431-447) */
typedef struct {
  double kp_position, kd_position;            /* before the controller is loaded                     */
  double kp_big_stance, kp_big_swing, kd_big; /* hip pitch / knee (joints 2,3,7,8)                   */
  double kp_small_stance, kp_small_swing, kd_small; /* hip roll / yaw (0,1,5,6) and ankle kp (4,9)   */
  double kd_feet;                             /* ankle (4,9)                                         */
} hb_pd_gains;

/* ---- state estimator: linear Kalman filter of legged_estimation/src/LinearKalmanFilter.cpp:72-185 ---- */
typedef struct {
  double x_hat[18];        /* base position(3), base linear velocity(3), four contact positions(12), world frame */
  double P[18 * 18];       /* covariance, row-major                                                              */
  double feet_heights[4];  /* terrain height under each contact (measurement rows 24..27)                        */
} hb_kf_state;

typedef struct {           /* task.info:336-345 */
  double foot_radius, imu_process_noise_position, imu_process_noise_velocity, foot_process_noise_position;
  double foot_sensor_noise_position, foot_sensor_noise_velocity, foot_height_sensor_noise;
} hb_kf_params;

/* state of the generalised-momentum observer behind the contact-force estimate (pSCgZinvlast_, StateEstimateBase.h:125) */
typedef struct { double p_filtered[16]; } hb_observer_state;
int hb_observer_reset(int B, hb_observer_state* state);   /* host only: zeros (StateEstimateBase.cpp:58-59) */

/* ---- run-time WBC settings: what WbcBase::loadTasksSetting / WeightedWbc::loadTasksSetting read from task.info
 * (legged_wbc/src/WbcBase.cpp:352-411, WeightedWbc.cpp:96-111) and WbcBase::setKpKd changes (WbcBase.h:65-69) ---- */
typedef struct {
  double torque_limits[5];            /* torqueLimitsTask (motor 1..5 of a leg)                    task.info:290-297 */
  double friction_coefficient;        /* frictionConeTask.frictionCoefficient                      :299-302          */
  double swing_kp, swing_kd;          /* swingLegTask                                              :304-308          */
  double base_accel_kp, base_accel_kd;   /* baseAccelTask (loaded by the reference, used by no task) :310-314        */
  double base_height_kp, base_height_kd; /* baseHeightTask                                          :316-320         */
  double base_angular_kp, base_angular_kd; /* baseAngularTask                                       :322-326         */
  double weight_swing_leg, weight_base_accel, weight_contact_force;   /* weight                     :328-333         */
} hb_wbc_settings;

/* what hb_parse_task_info extracts from a task.info file (boost property-tree INFO format): the WBC block above, the estimator
 * parameters and the solver discretisation. `found` has one bit per section that was present (1 WBC, 2 kalmanFilter,
 * 4 contactForceEsimation, 8 sqp, 16 mpc); absent sections keep the defaults. */
typedef struct {
  hb_wbc_settings wbc;
  double kalman[7];                   /* hb_kf_params in declaration order (kalmanFilter, task.info:336-345)                */
  double contact_force_cutoff_frequency, contact_threshold;   /* contactForceEsimation                   :347-351          */
  double sqp_dt;                      /* sqp.dt                                                            :82               */
  int32_t sqp_iteration;              /* sqp.sqpIteration                                                  :83               */
  double mpc_time_horizon;            /* mpc.timeHorizon                                                   :144              */
  int32_t mpc_cold_start;             /* mpc.coldStart                                                     :146              */
  int32_t found;
} hb_task_info;

int hb_default_wbc_settings(hb_wbc_settings* s);
int hb_parse_task_info(const char* path, hb_task_info* out);          /* host only; -1: file missing or malformed */
/* the WBC settings in force for this context (defaults: the shipped task.info values compiled into include/hunter_model_constants.h) */
int hb_wbc_get_settings(const hb_ctx* ctx, hb_wbc_settings* s);
int hb_wbc_set_settings(hb_ctx* ctx, const hb_wbc_settings* s);
int hb_wbc_set_kp_kd(hb_ctx* ctx, double swing_kp, double swing_kd);   /* WbcBase::setKpKd */
int hb_load_task_info(hb_ctx* ctx, const char* path);                  /* hb_parse_task_info + hb_wbc_set_settings */

/* ---- hierarchical QP: legged::HoQp / legged::HierarchicalWbc ---- */
#define HB_HOQP_MAX_LEVELS 3
#define HB_HOQP_N 38            /* decision variables (the WBC's [qdd, F, tau])                             */
#define HB_HOQP_MAX_EQ 32       /* rows of a (equality task) per level                                      */
#define HB_HOQP_MAX_IN 40       /* rows of d (inequality task) per level                                    */
#define HB_HOQP_MAX_STACKED 80  /* inequality rows of all levels together                                   */
typedef struct {                /* one hierarchy: level 0 has the highest priority (Task a x = b, d x <= f; legged_wbc/include/legged_wbc/Task.h) */
  int32_t n, levels;
  int32_t ma[HB_HOQP_MAX_LEVELS], md[HB_HOQP_MAX_LEVELS];
  double a[HB_HOQP_MAX_LEVELS][HB_HOQP_MAX_EQ][HB_HOQP_N], b[HB_HOQP_MAX_LEVELS][HB_HOQP_MAX_EQ];
  double d[HB_HOQP_MAX_LEVELS][HB_HOQP_MAX_IN][HB_HOQP_N], f[HB_HOQP_MAX_LEVELS][HB_HOQP_MAX_IN];
} hb_hoqp_problem;

/* ---- closed-loop rollout: actuation model of the simulated hardware and a batched rigid-body plant ---- */
#define HB_ACT_CAPACITY 16
typedef struct {           /* command buffer of one robot (LeggedHWSim::cmdBuffer_, legged_gazebo/src/LeggedHWSim.cpp:166-186) */
  int32_t count, head;     /* entries in the ring; index of the newest one                                  */
  double stamp[HB_ACT_CAPACITY];
  double cmd[HB_ACT_CAPACITY][50];   /* per joint: posDes, velDes, kp, kd, ff                               */
} hb_actuation_state;
typedef struct {
  double dt;                /* control period covered by one call [s] (500 Hz loop: 0.002)                  */
  int32_t substeps;         /* semi-implicit Euler substeps per call                                        */
  double ground_height;     /* flat ground z                                                                */
  double ground_stiffness, ground_damping;   /* normal spring-damper per contact point [N/m], [N s/m]       */
  double tangential_damping;                 /* viscous tangential friction [N s/m], clipped to mu * F_z    */
  double friction_mu;
  double joint_armature;    /* rotor inertia added to the joint diagonal of M [kg m^2] (mujoco/model/hunter/hunter.xml:6: 0.1) */
  double joint_damping;     /* viscous joint damping [N m s/rad] (hunter.xml:6: 1)                           */
} hb_sim_params;
int hb_default_sim_params(hb_sim_params* p);
int hb_actuation_reset(int B, hb_actuation_state* state);   /* host only */

int hb_default_kf_params(hb_kf_params* p);
/* x_hat = 0, P = 100 I, heights = 0 (KalmanFilterEstimate constructor, LinearKalmanFilter.cpp:24-63); host only */
int hb_kf_reset(int B, hb_kf_state* state);
int hb_default_pd_gains(hb_pd_gains* g);
int hb_default_config(hb_config* cfg);
int hb_create(const hb_config* cfg, int device, hb_ctx** out);
int hb_destroy(hb_ctx* ctx);
int hb_sync(hb_ctx* ctx);
const char* hb_strerror(int code);
/* text of the last CUDA runtime error seen by this context (diagnostics for return code -2) */
const char* hb_last_cuda_error(const hb_ctx* ctx);
/* number of kernel launches issued through this context since creation (bench.py reports it as gpu_launches) */
int64_t hb_launch_count(const hb_ctx* ctx);
/* bytes of reference data the last hb_resident_cycle_batch moved host -> device (only the used entries of hb_reference are uploaded) */
int64_t hb_last_reference_upload_bytes(const hb_ctx* ctx);
/* per-kernel device timing with CUDA events on the context's stream (used by bench.py for the roofline line):
 * kinds 0 = Riccati sweep, 1 = forward pass + line search, 2 = WBC assembly, 3 = interior-point QP, 4 = other,
 *       5 = node linearisation (kinematics), 6 = node LQ model + projection */
int hb_profile_enable(hb_ctx* ctx, int on);
int hb_profile_read(hb_ctx* ctx, double* ms_per_kind /*7*/, int64_t* count_per_kind /*7*/);
/* the context's CUDA stream as a cudaStream_t cast to void* (for event timing on the launching stream) */
void* hb_stream(hb_ctx* ctx);

/* ---- device-pointer (asynchronous) entry points ---- */
int hb_wbc_qp_batch_dev(hb_ctx* ctx, int B, int n, int m, const double* H, const double* g, const double* A, const double* lbA,
                        const double* ubA, double* x, int32_t* status, int32_t* iters);
int hb_wbc_solve_batch_dev(hb_ctx* ctx, int B, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                           const uint8_t* stance_mode, double* sol, int32_t* status);
/* WbcBase::formulate*Task + WeightedWbc::formulateConstraints / formulateWeightedTasks (legged_wbc/src/WbcBase.cpp:138-338,
 * WeightedWbc.cpp:68-94) in the layout WeightedWbc::update hands to qpOASES (WeightedWbc.cpp:24-42): H = A_w' A_w (38 x 38), g = -A_w' b_w,
 * A (60 rows allocated, m_rows[i] used, row-major 60 x 38), lbA / ubA (60, -1e20 = qpOASES -INFTY). Feeds the raw QP sweep
 * (BASELINE configs[4]) and the assembly parity test; the product path (hb_wbc_solve_batch) never materialises these matrices. */
int hb_wbc_assemble_batch_dev(hb_ctx* ctx, int B, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                              const uint8_t* stance_mode /*nullable*/, double* H, double* g, double* A, double* lbA, double* ubA,
                              int32_t* m_rows);
/* hb_wbc_qp_batch_dev with a per-problem row count: A / lbA / ubA are allocated with m_alloc rows, problem i uses its first m_rows[i] */
int hb_wbc_qp_rows_batch_dev(hb_ctx* ctx, int B, int n, int m_alloc, const int32_t* m_rows, const double* H, const double* g,
                             const double* A, const double* lbA, const double* ubA, double* x, int32_t* status, int32_t* iters);
/* HoQp cascade (legged_wbc/src/HoQp.cpp:21-198): x (B x 38, first n entries used) = solution of the lowest level (HoQp::getSolutions),
 * slack (B x 80, nullable) = stacked slack solutions in level order, status[i] = 0 or 10 * (QP status) + level of the first failing level */
int hb_hoqp_solve_batch_dev(hb_ctx* ctx, int B, const hb_hoqp_problem* problems, double* x, double* slack /*nullable*/, int32_t* status /*nullable*/);
/* legged::HierarchicalWbc::update (legged_wbc/src/HierarchicalWbc.cpp:18-31): task0 = floating-base EoM + torque limits + friction cone +
 * no contact motion, task1 = base acceleration, task2 = 0.1 * contact force + swing leg; sol (B x 38) = [qdd, F, tau] */
int hb_hierarchical_wbc_solve_batch_dev(hb_ctx* ctx, int B, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                                        double* sol, int32_t* status /*nullable*/);
int hb_mpc_cold_start_batch_dev(hb_ctx* ctx, int B, const double* x0, const int32_t* mode, double* x_traj, double* u_traj);
int hb_mpc_solve_batch_dev(hb_ctx* ctx, int B, const double* x0, const double* x_ref, const double* swing_ref, const int32_t* mode,
                           double* x_traj, double* u_traj, hb_solve_info* info);
/* ---- time discretisation with event nodes ----
 * hb_time_grid_batch: node times (B x (horizon_N + 1)) and interval counts (B) of every instance from its mode-switch times
 * (hb_reference.event_times); status[i] = 1 when the capacity horizon_N was exhausted (last interval stretched to the final time).
 * The *_grid_* entry points are the grid-aware forms of their uniform counterparts: interval k of instance i has length
 * node_times[i][k+1] - node_times[i][k]; nodes beyond n_intervals[i] are ignored. */
int hb_time_grid_batch_dev(hb_ctx* ctx, int B, const double* t0, const hb_reference* refs, double* node_times, int32_t* n_intervals,
                           int32_t* status /*nullable*/);
int hb_reference_expand_grid_batch_dev(hb_ctx* ctx, int B, const double* node_times, const hb_reference* refs, double* x_ref, double* swing_ref,
                                       int32_t* mode);
int hb_mpc_solve_grid_batch_dev(hb_ctx* ctx, int B, const double* x0, const double* node_times, const int32_t* n_intervals, const double* x_ref,
                                const double* swing_ref, const int32_t* mode, double* x_traj, double* u_traj, hb_solve_info* info);
int hb_policy_eval_grid_batch_dev(hb_ctx* ctx, int B, double t_rel, const double* node_times, const int32_t* n_intervals, const double* x_traj,
                                  const double* u_traj, const int32_t* mode, double* x_des, double* u_des, int32_t* mode_out);
int hb_policy_eval_batch_dev(hb_ctx* ctx, int B, double t_rel, const double* x_traj, const double* u_traj, const int32_t* mode,
                             double* x_des, double* u_des, int32_t* mode_out);
int hb_control_step_batch_dev(hb_ctx* ctx, int B, double t_rel, const double* x0, const double* x_ref, const double* swing_ref,
                              const int32_t* mode, const double* rbd, double* x_traj, double* u_traj, hb_solve_info* info,
                              double* wbc_sol, double* torque, int32_t* wbc_status);
/* joint command law (LeggedController.cpp:186-257): posDes = q_mpc + 0.5 qdd dt^2, velDes = qd_mpc + qdd dt, gains by joint class
 * and planned contact state of the leg, feed-forward = WBC torque; joint-limit protection sets the per-instance emergency stop flag
 * (in/out) which turns the command into pure damping (0,0,0,1,0). loaded[i] = 0 selects the pre-load position hold (:211-222).
 * command: B x 10 x 5 = (posDes, velDes, kp, kd, ff) per joint; output_torque: B x 10 = ff + kp (posDes - q) + kd (velDes - qd). */
int hb_joint_command_batch_dev(hb_ctx* ctx, int B, const hb_pd_gains* gains, double period, const double* x_des, const double* u_des,
                               const double* wbc_sol, const int32_t* mode_cmd, const double* rbd, const uint8_t* loaded, uint8_t* estop,
                               double* command, double* output_torque);
/* Resident closed-loop cycle: the context keeps the primal solution on the device like ocs2::SqpSolver keeps primalSolution_.
 * One call = reference expansion (hb_reference -> node grid) + warm start (previous solution interpolated on the new grid, tail from
 * the initializer; mpc.coldStart false, task.info:146 -- or the initializer everywhere when cold_start != 0) + one SQP iteration +
 * policy evaluation at t0 + t_rel + WeightedWbc + torque law. Only t0, x0, refs, rbd go in and info / wbc_sol / torque / status come
 * out; hb_resident_read_batch copies the resident trajectories out when the caller wants them (PrimalSolution). The WeightedWbc
 * fallback is applied on the device: an instance whose QP did not solve (wbc_status != 0) returns its previous solution and torques
 * (WeightedWbc.cpp:57-64) from the second cycle on. */
int hb_resident_cycle_batch_dev(hb_ctx* ctx, int B, int cold_start, double t_rel, const double* t0, const double* x0,
                                const hb_reference* refs, const double* rbd, hb_solve_info* info, double* wbc_sol, double* torque,
                                int32_t* wbc_status);
/* device planner: the same planner source as hb_plan_references, four cooperating threads per instance. feet != NULL
 * overrides in[i].feet_pos (B x 12, e.g. from hb_contact_positions_batch_dev); status[i] = 0, -1 or -5 like hb_plan_references. */
int hb_plan_references_batch_dev(hb_ctx* ctx, int B, const hb_plan_input* in, const double* feet, double* latest_stance,
                                 hb_reference* out, int32_t* status /*nullable*/);
/* One estimator update per instance (StateEstimateBase::updateJointStates / updateImu, StateEstimateBase.cpp:73-106, then
 * KalmanFilterEstimate::update): quat = (x, y, z, w); contact_flag: B x 4 (0 = the filter distrusts that foot, x100 noise);
 * rbd_out: B x 32 measured rbd state [zyx, p, q_j, omega_world, v, qd_j]. The odometry topic fusion (updateFromTopic) is ROS glue
 * and not part of it; zyxOffset_ is taken as zero. */
int hb_estimator_update_batch_dev(hb_ctx* ctx, int B, const hb_kf_params* params, double dt, hb_kf_state* state, const double* quat,
                                  const double* ang_vel_local, const double* lin_acc_local, const double* joint_pos,
                                  const double* joint_vel, const uint8_t* contact_flag, double* rbd_out);
/* StateEstimateBase::estContactForce (legged_estimation/src/StateEstimateBase.cpp:130-206): momentum-observer disturbance torque and the
 * least-norm 6-D wrench at the toe frame of each foot. cutoff_frequency = contactForceEsimation.cutoffFrequency (task.info:349, 250);
 * dt > 1 is replaced by 0.002 as in the reference. rbd: measured state (B x 32), tau_cmd: last commanded joint torques (B x 10).
 * est_contact_force (B x 16) = [wrench_left(6), wrench_right(6), |F_l|, |F_r|, |W_l|, |W_r|]; disturbance_torque (B x 16, nullable). */
int hb_contact_force_estimate_batch_dev(hb_ctx* ctx, int B, double cutoff_frequency, double dt, hb_observer_state* state, const double* rbd,
                                        const double* tau_cmd, double* est_contact_force, double* disturbance_torque /*nullable*/);
/* LeggedHWSim::writeSim (legged_gazebo/src/LeggedHWSim.cpp:166-192): command (B x 10 x 5 as written by hb_joint_command_batch) delayed by up to
 * `delay` seconds (legged_gazebo/config/default.yaml:2, 0.009), PD + feed-forward evaluated with the current joint state -> tau (B x 10) */
int hb_actuation_batch_dev(hb_ctx* ctx, int B, double delay, const double* time /*B*/, hb_actuation_state* state, const double* command,
                           const double* rbd, double* tau);
/* one control period of the batched plant: rbd (B x 32, [zyx, p, q_j, omega_world, v, qd_j]) advanced in place under the joint torques tau
 * (B x 10); contact_force (B x 12) and contact_flag (B x 4) of the last substep are optional outputs */
int hb_sim_step_batch_dev(hb_ctx* ctx, int B, const hb_sim_params* params, double* rbd, const double* tau, double* contact_force /*nullable*/,
                          uint8_t* contact_flag /*nullable*/);
/* The 500 Hz half of LeggedController::update between two MPC solves (LeggedController.cpp:154-184): evaluatePolicy of the RESIDENT solution at
 * the absolute time t_now[i], WeightedWbc (with the previous-solution fallback), torque law. Outputs the desired state / input / mode too
 * (inputs of hb_joint_command_batch). */
int hb_resident_wbc_batch_dev(hb_ctx* ctx, int B, const double* t_now, const double* rbd, const uint8_t* stance_mode /*nullable*/, double* x_des,
                              double* u_des, int32_t* mode_out, double* wbc_sol, double* torque /*nullable*/, int32_t* wbc_status /*nullable*/);
int hb_rbd_to_centroidal_batch_dev(hb_ctx* ctx, int B, const double* rbd, double* x);
int hb_reference_expand_batch_dev(hb_ctx* ctx, int B, const double* t0, const hb_reference* refs, double* x_ref, double* swing_ref,
                                  int32_t* mode);
/* world positions of the four contact frames at the configuration of x (InverseKinematics::computeFootPos,
 * legged_interface/src/foot_planner/InverseKinematics.cpp:253-267) */
int hb_contact_positions_batch_dev(hb_ctx* ctx, int B, const double* x, double* pos /*B x 12*/);
/* probe used by the parity tests: the shipping node linearisation (lin_half of K0) expanded to full tiles: f (22), A = df/dx and
 * Bm = df/du (22 x 22), ee = [pos(12), vel(12), dpos/dx (12 x 22), dvel/dx (12 x 22), dvel/du (12 x 22)] */
int hb_probe_flow_map_dev(hb_ctx* ctx, int B, const double* x, const double* u, double* f, double* A, double* Bm, double* ee);

/* ---- host-pointer (synchronous) entry points: H2D copy, device call, D2H copy ---- */
int hb_wbc_qp_batch(hb_ctx* ctx, int B, int n, int m, const double* H, const double* g, const double* A, const double* lbA,
                    const double* ubA, double* x, int32_t* status, int32_t* iters);
int hb_wbc_solve_batch(hb_ctx* ctx, int B, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                       const uint8_t* stance_mode, double* sol, int32_t* status);
int hb_wbc_assemble_batch(hb_ctx* ctx, int B, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                          const uint8_t* stance_mode /*nullable*/, double* H, double* g, double* A, double* lbA, double* ubA, int32_t* m_rows);
int hb_hoqp_solve_batch(hb_ctx* ctx, int B, const hb_hoqp_problem* problems, double* x, double* slack /*nullable*/, int32_t* status /*nullable*/);
int hb_hierarchical_wbc_solve_batch(hb_ctx* ctx, int B, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode, double* sol,
                                    int32_t* status /*nullable*/);
/* the three tasks HierarchicalWbc::update builds, one hb_hoqp_problem per instance; for inspection / parity tests */
int hb_hierarchical_wbc_tasks_batch(hb_ctx* ctx, int B, const double* x_des, const double* u_des, const double* rbd, const int32_t* mode,
                                    hb_hoqp_problem* problems);
int hb_mpc_cold_start_batch(hb_ctx* ctx, int B, const double* x0, const int32_t* mode, double* x_traj, double* u_traj);
int hb_mpc_solve_batch(hb_ctx* ctx, int B, const double* x0, const double* x_ref, const double* swing_ref, const int32_t* mode,
                       double* x_traj, double* u_traj, hb_solve_info* info);
int hb_control_step_batch(hb_ctx* ctx, int B, double t_rel, const double* x0, const double* x_ref, const double* swing_ref,
                          const int32_t* mode, const double* rbd, double* x_traj, double* u_traj, hb_solve_info* info,
                          double* wbc_sol, double* torque, int32_t* wbc_status);
int hb_joint_command_batch(hb_ctx* ctx, int B, const hb_pd_gains* gains, double period, const double* x_des, const double* u_des,
                           const double* wbc_sol, const int32_t* mode_cmd, const double* rbd, const uint8_t* loaded, uint8_t* estop,
                           double* command, double* output_torque);
/* Host-pointer resident cycle. Only the USED entries of the fixed-capacity hb_reference structs cross PCIe. If `refs` points into page-locked
 * memory (cudaHostAlloc / cudaHostRegister; the whole array refs[0..B) must lie inside that allocation) the device gathers and validates the
 * entries itself and the call does no per-instance host work; a pageable array is validated and packed on the host first. Malformed structs
 * (counts beyond the capacities, unordered times, an empty target list, modes outside 0..3) return -1 on both paths. */
int hb_resident_cycle_batch(hb_ctx* ctx, int B, int cold_start, double t_rel, const double* t0, const double* x0, const hb_reference* refs,
                            const double* rbd, hb_solve_info* info, double* wbc_sol, double* torque, int32_t* wbc_status);
int hb_resident_read_batch(hb_ctx* ctx, int B, double* t0 /*nullable*/, double* x_traj /*nullable*/, double* u_traj /*nullable*/);
/* Restores a snapshot taken with hb_resident_read_batch (+ hb_resident_read_grid_batch): the next warm cycle shifts it exactly as if this
 * context had produced it (checkpoint / resume, migration of instances between contexts or GPUs). mode (B x (N+1), nullable) = node modes of
 * the solution, needed by hb_resident_wbc_batch until the next cycle; node_times / n_intervals are required by event_nodes contexts only. */
int hb_resident_write_batch(hb_ctx* ctx, int B, const double* t0, const double* x_traj, const double* u_traj, const int32_t* mode /*nullable*/,
                            const double* node_times /*nullable*/, const int32_t* n_intervals /*nullable*/);
/* node times / interval counts of the resident solution (contexts created with event_nodes = 1) */
int hb_resident_read_grid_batch(hb_ctx* ctx, int B, double* node_times, int32_t* n_intervals);
int hb_time_grid_batch(hb_ctx* ctx, int B, const double* t0, const hb_reference* refs, double* node_times, int32_t* n_intervals, int32_t* status /*nullable*/);
int hb_reference_expand_grid_batch(hb_ctx* ctx, int B, const double* node_times, const hb_reference* refs, double* x_ref, double* swing_ref, int32_t* mode);
int hb_mpc_solve_grid_batch(hb_ctx* ctx, int B, const double* x0, const double* node_times, const int32_t* n_intervals, const double* x_ref,
                            const double* swing_ref, const int32_t* mode, double* x_traj, double* u_traj, hb_solve_info* info);
/* hb_plan_references on the device, host pointers in and out (parity checks of the device planner against the host planner) */
int hb_plan_references_gpu(hb_ctx* ctx, int B, const hb_plan_input* in, double* latest_stance, hb_reference* out, int32_t* status /*nullable*/);
/* The whole cycle from the plan inputs: computeFootPos at x0 + planner (P1, P3, P4, P5) + hb_resident_cycle_batch, all on the device.
 * in[i].feet_pos is ignored (computed from in[i].x0); the planner's latest-stance state is resident (zeroed by cold_start, like
 * SwingTrajectoryPlanner's latestStanceposition_). Per instance 352 B + rbd go in, info / wbc_sol / torque / statuses come out.
 * plan_status[i] != 0: the planner rejected the instance (an all-stance reference at the current pose was used instead). */
int hb_resident_plan_cycle_batch(hb_ctx* ctx, int B, int cold_start, double t_rel, const hb_plan_input* in, const double* rbd,
                                 hb_solve_info* info, double* wbc_sol, double* torque, int32_t* wbc_status, int32_t* plan_status);
int hb_estimator_update_batch(hb_ctx* ctx, int B, const hb_kf_params* params, double dt, hb_kf_state* state, const double* quat,
                              const double* ang_vel_local, const double* lin_acc_local, const double* joint_pos, const double* joint_vel,
                              const uint8_t* contact_flag, double* rbd_out);
int hb_contact_force_estimate_batch(hb_ctx* ctx, int B, double cutoff_frequency, double dt, hb_observer_state* state, const double* rbd,
                                    const double* tau_cmd, double* est_contact_force, double* disturbance_torque /*nullable*/);
int hb_actuation_batch(hb_ctx* ctx, int B, double delay, const double* time, hb_actuation_state* state, const double* command, const double* rbd,
                       double* tau);
int hb_sim_step_batch(hb_ctx* ctx, int B, const hb_sim_params* params, double* rbd, const double* tau, double* contact_force /*nullable*/,
                      uint8_t* contact_flag /*nullable*/);
int hb_resident_wbc_batch(hb_ctx* ctx, int B, const double* t_now, const double* rbd, const uint8_t* stance_mode /*nullable*/, double* x_des, double* u_des,
                          int32_t* mode_out, double* wbc_sol, double* torque /*nullable*/, int32_t* wbc_status /*nullable*/);
int hb_rbd_to_centroidal_batch(hb_ctx* ctx, int B, const double* rbd, double* x);
int hb_reference_expand_batch(hb_ctx* ctx, int B, const double* t0, const hb_reference* refs, double* x_ref, double* swing_ref,
                              int32_t* mode);
int hb_probe_flow_map(hb_ctx* ctx, int B, const double* x, const double* u, double* f, double* A, double* Bm, double* ee);
int hb_contact_positions_batch(hb_ctx* ctx, int B, const double* x, double* pos /*B x 12*/);

/* ---- host-only reference preprocessing (no GPU work): gait tiling, swing-foot planner, cmd_vel target ----
 * replaces GaitSchedule::{insert,tile}ModeSequenceTemplate (legged_interface/src/gait/GaitSchedule.cpp:57-161),
 * SwingTrajectoryPlanner::update (src/foot_planner/SwingTrajectoryPlanner.cpp:164-286), cmdVelToTargetTrajectories
 * (legged_controllers/src/TargetTrajectoriesPublisher.cpp:102-130) and calculateJointRef + InverseKinematics::computeIK
 * (src/SwitchedModelReferenceManager.cpp:251-300, src/foot_planner/InverseKinematics.cpp:20-231). latest_stance (B x 12) is the planner's state, in/out.
 * Returns 0, or -1 on misuse, or -5 when a schedule does not define the take-off / touch-down of a swing phase (the reference
 * throws there, SwingTrajectoryPlanner.cpp:421-458) or exceeds the capacity of hb_reference.
 * Instances are spread over std::thread::hardware_concurrency() host threads (hb_plan_set_threads overrides the count). */
int hb_plan_references(int B, const hb_plan_input* in, double* latest_stance, hb_reference* out);
/* Host threads used by hb_plan_references (process-wide); 0 = hardware_concurrency. The result does not depend on the count. */
int hb_plan_set_threads(int n_threads);
/* speed-based gait selection (calculateVelAbs + walkGait / trotGait, src/SwitchedModelReferenceManager.cpp:185-249): updates the
 * 50-sample moving average of 0.5*(command + target) speed of every instance and applies the thresholds stance <= 0.02 < (no change)
 * <= 0.03 < trot < 0.4 <= level 3. gait_type: 0 walk (automatic), 2 trot (forced). level[i] = gait level in force after the call,
 * insert[i] = 1 when a new template is inserted on this call. */
int hb_gait_select(int B, hb_gait_selector* state, const int32_t* gait_type, const double* cmd_vel /*B x 4*/,
                   const double* target_state0 /*B x 22*/, int32_t* level, int32_t* insert);

/* ---- (e) multi-GPU shards (replaces nothing in the reference, which runs one robot per process): one process + one hb_ctx
 * per GPU, instances split in contiguous blocks, no collective on the data path. The only exchange is the gather of per-instance
 * output rows (the 80-byte torque rows of the control step), by NCCL all-gather on the shard's own stream behind an event on the
 * context's stream, so it overlaps the next step's kernels. NCCL is resolved at run time (the libnccl.so.2 already in the process, else
 * the system one); a single-GPU caller never loads it. Return code -6: NCCL missing or a collective failed (hb_shard_last_error). */
#define HB_SHARD_ID_BYTES 128
typedef struct hb_shard hb_shard;
/* contiguous block [begin, begin + count) of `rank`; block sizes differ by at most one */
int hb_shard_partition(int total, int world, int rank, int* begin, int* count);
/* stable permutation that groups instances with the same mode sequence (mode: B x nodes), so that the warps of a wave run the same
 * swing-contact specialisation: sorted[i] = original[perm[i]]; inverse (nullable): inverse[perm[i]] = i */
int hb_shard_sort_by_schedule(int B, int nodes, const int32_t* mode, int32_t* perm, int32_t* inverse);
/* rank 0 draws the communicator id (HB_SHARD_ID_BYTES bytes); the caller hands the bytes to the other ranks out of band (MPI, a file, a socket) */
int hb_shard_unique_id(void* id);
/* collective over all ranks; total_instances = instances of the whole job, max_row_doubles = widest row ever gathered. world == 1: id may be NULL */
int hb_shard_create(hb_ctx* ctx, const void* id, int world, int rank, int total_instances, int max_row_doubles, hb_shard** out);
int hb_shard_destroy(hb_shard* shard);
int hb_shard_block(const hb_shard* shard, int* begin, int* count);
/* rows_dev: count x row_doubles of this rank's block (device). inverse_dev (nullable, device): undo a schedule sort first
 * (row i of the block = rows_dev[inverse_dev[i]]). *gathered_dev: total_instances x row_doubles in instance order on EVERY rank, complete
 * after hb_shard_wait; double buffered, valid until the second next call. Asynchronous: ordered behind the context's stream. */
int hb_shard_gather_dev(hb_shard* shard, int row_doubles, const double* rows_dev, const int32_t* inverse_dev, const double** gathered_dev);
/* the context's stream waits for the last gather; block_host != 0 also blocks the calling thread until it has landed */
int hb_shard_wait(hb_shard* shard, int block_host);
const char* hb_shard_last_error(const hb_shard* shard);

#ifdef __cplusplus
}
#endif
#endif
