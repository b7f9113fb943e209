"""Mirror of the entry points of ``klampt.plan.robotplanning`` that sit directly on the feasibility path (reference
Python/klampt/plan/robotplanning.py:23-265): ``make_space`` builds the robot's C-space for a world (collider with ignored pairs,
extra constraints, edge resolution, optional moving subset) and ``plan_to_config`` sets up a planner from the robot's current
configuration to a target -- here a batched planner (klampt_b200.plan.MotionPlan) over the GPU engine.

Robots with affine drivers get the driver-space embedding (AffineEmbeddedCSpace), as in the reference.  Not mirrored: equality
constraints (closed-loop / IK spaces), which raise NotImplementedError.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Union

from . import collide
from .cspaceutils import AffineEmbeddedCSpace, EmbeddedCSpace, EmbeddedMotionPlan
from .plan import MotionPlan


def active_links(robot, subset: Sequence[int]) -> List[bool]:
    """links that can move when only the DOFs in `subset` do: the subset and everything below it (reference
    robotcspace.py:372-377).  sic: a root link looks up ``active[-1]`` there, i.e. the flag of the LAST link at that moment, so a root
    counts as active whenever the last link is in the subset; kept, it only leaves more pairs enabled."""
    n = robot.numLinks()
    active = [False] * n
    for i in subset:
        active[i] = True
    for i in range(n):
        if active[robot.link(i).getParent()]:
            active[i] = True
    return active


def disable_inactive_collisions(collider: collide.WorldCollider, robot, subset: Sequence[int]) -> None:
    """EmbeddedRobotCSpace.disableInactiveCollisions (reference robotcspace.py:365-392): a link that cannot move is only checked
    against links that can -- its pairs with the environment and with other fixed links are constant and are dropped from the mask"""
    active = active_links(robot, subset)
    rindices = collider.robots[robot.index]
    for i in range(robot.numLinks()):
        if active[i] or rindices[i] < 0:
            continue
        collider.mask[rindices[i]] = {rindices[j] for j in range(robot.numLinks()) if rindices[j] in collider.mask[rindices[i]] and active[j]}


def make_space(world, robot, edgeCheckResolution: float = 1e-2, extraConstraints: Sequence[Callable] = (), equalityConstraints: Sequence = (),
               equalityTolerance: float = 1e-3, ignoreCollisions: Sequence = (), movingSubset: Optional[Union[str, Sequence[int]]] = None, device: int = 0):
    """the C-space of `robot` in `world`; with a moving subset, an EmbeddedCSpace over it whose fixed DOFs stay at the robot's current
    configuration.  The engine is built from the collider's mask AFTER the inactive pairs are dropped, so they cost nothing."""
    from .robotcspace import RobotCSpace
    if len(equalityConstraints) > 0:
        raise NotImplementedError("equality (closed-loop) constraints are not supported by the batched space")
    subset = None if movingSubset in ("auto", "all", None) else list(movingSubset)
    collider = collide.WorldCollider(world, ignore=list(ignoreCollisions))
    embedded = subset is not None and len(subset) < robot.numLinks()
    if embedded:
        disable_inactive_collisions(collider, robot, subset)
    space = RobotCSpace(robot, collider, device=device)
    space.eps = edgeCheckResolution
    for c in extraConstraints:
        space.addConstraint(c)
    # robots with affine drivers (coupled links: mimic joints, grippers) are planned in driver space (reference :94-142)
    drivers = list(getattr(getattr(getattr(space, "spec", None), "robot", None), "drivers", []) or [])
    moving = set(range(robot.numLinks()) if subset is None else subset)
    if any(len(d.links) > 1 and moving.intersection(d.links) for d in drivers):
        active = [k for k, d in enumerate(drivers) if moving.intersection(d.links)]
        aff = AffineEmbeddedCSpace.from_drivers(space, drivers, robot.numLinks(), active)
        aff.robot = robot
        aff.setup()
        return aff
    if embedded:
        space = EmbeddedCSpace(space, subset, xinit=robot.getConfig())
        space.robot = robot
    space.setup()
    return space


def plan_to_config(world, robot, target: Sequence[float], edgeCheckResolution: float = 1e-2, extraConstraints: Sequence[Callable] = (),
                   equalityConstraints: Sequence = (), equalityTolerance: float = 1e-3, ignoreCollisions: Sequence = (),
                   movingSubset: Optional[Union[str, Sequence[int]]] = "auto", verbose: bool = True, device: int = 0, **planOptions):
    """a planner from the robot's current configuration to `target` (reference robotplanning.py:153-265): 'auto' moves exactly the
    DOFs whose start and target values differ; a fixed DOF whose values differ is an error; returns None (with a warning naming the
    failing tests) when an endpoint is infeasible"""
    import warnings
    q0 = robot.getConfig()
    if len(q0) != len(target):
        raise ValueError("target configuration must be of correct size for robot")
    if movingSubset == "auto":
        subset = [i for i, (a, b) in enumerate(zip(q0, target)) if a != b]
    elif movingSubset == "all" or movingSubset is None:
        subset = list(range(len(q0)))
    else:
        subset = list(movingSubset)
        for i in range(len(q0)):
            if i not in subset and q0[i] != target[i]:
                raise ValueError("Error: target configuration value differs from start configuration along a fixed DOF: %s (link %d): %g vs %g"
                                 % (robot.link(i).getName(), i, q0[i], target[i]))
    space = make_space(world, robot, edgeCheckResolution, extraConstraints, equalityConstraints, equalityTolerance, ignoreCollisions, subset, device)
    plan = EmbeddedMotionPlan(space, q0, **planOptions) if hasattr(space, "lift") else MotionPlan(space, **planOptions)
    try:
        plan.setEndpoints(q0, list(target))
    except RuntimeError:
        amb = getattr(space, "ambientspace", space)
        for name, q in (("Start", q0), ("Goal", list(target))):
            fails = amb.feasibilityFailures(list(q))
            if fails and verbose:
                warnings.warn("%s configuration fails %s" % (name, fails))
        return None
    return plan
