"""hunter_bipedal_control_b200 -- B200-native batched NMPC + WBC solve path for the Hunter biped (one hot path, not a port).

Only what the path needs: ``csrc/`` (sm_100a CUDA kernels + the C ABI of include/hunter_b200.h) and ``api`` (ctypes binding and
host-side mirrors of the reference's WbcBase / MPC_MRT_Interface calls). No CPU fallback exists.
"""
from .api import (HbWbcSettings, HbTaskInfo, parse_task_info, Context, WeightedWbc, HierarchicalWbc, HbHoqpProblem, make_hoqp_problems, hoqp_tasks, SqpMpc, HbReference, HbSolveInfo, HbConfig, HunterB200Error, load_library, EXPORTED_SYMBOLS,
                  NX, NU, NQ, NJ, NWBC, INFO_DTYPE, HbPlanInput, plan_references, plan_set_threads, make_plan_inputs, GAIT_IDS, GaitSelector, HbPdGains, default_pd_gains, HbKfState, HbKfParams, default_kf_params, kf_states, HbObserverState, observer_states, HbActuationState, HbSimParams, default_sim_params, actuation_states)

__all__ = ["HbWbcSettings", "HbTaskInfo", "parse_task_info", "Context", "WeightedWbc", "HierarchicalWbc", "HbHoqpProblem", "make_hoqp_problems", "hoqp_tasks", "SqpMpc", "HbReference", "HbSolveInfo", "HbConfig", "HunterB200Error", "load_library",
           "EXPORTED_SYMBOLS", "NX", "NU", "NQ", "NJ", "NWBC", "INFO_DTYPE", "HbPlanInput", "plan_references", "plan_set_threads", "make_plan_inputs", "GAIT_IDS", "GaitSelector", "HbPdGains", "default_pd_gains", "HbKfState", "HbKfParams", "default_kf_params", "kf_states", "HbObserverState", "observer_states", "HbActuationState", "HbSimParams", "default_sim_params", "actuation_states"]
