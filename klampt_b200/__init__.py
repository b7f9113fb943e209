"""klampt_b200 -- B200-native batched configuration-feasibility engine behind Klamp't's CSpace / robotsim collision
API: batched FK -> joint limits -> self + environment collision / distance -> discretised edge checks.

  klampt_b200.engine.Engine          the C ABI (include/klampt_b200.h) from Python; CUDA only, no CPU fallback
  klampt_b200.robotcspace.RobotCSpace  drop-in for klampt.plan.robotcspace.RobotCSpace + feasible_batch / visible_batch
  klampt_b200.robotsim / collide / cspace / so3   host-side mirrors of the reference interfaces on this path
  klampt_b200.plan / robotplanning / cspaceutils  batched planners (PRM, PRM*, Lazy-PRM*, RRT, SBL, shortcutting), make_space /
                                     plan_to_config, EmbeddedCSpace (planning on a DOF subset)
  klampt_b200.io                     OFF / OBJ / STL / PCD, .rob, URDF, world XML -> WorldSpec
  klampt_b200.distancequery          DistanceQuery's Far / Close / Contact cycle over batches of poses
  klampt_b200.synth                  seeded synthetic workloads C1..C5 (BASELINE.json configs)
  klampt_b200.shard                  one-process-per-GPU sharding and result gather
"""
__version__ = "0.1.0"
