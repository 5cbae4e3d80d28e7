"""redmax_b200 -- B200-native batched RedMax stepper behind the reference's +redmax Scene/Joint/Body API.

Host mirror of the object API (scene.py), scene factory (scenes.py), reference-named drivers (drivers.py) and the
ctypes binding of the C ABI (_ffi.py) over the CUDA library built from csrc/.  No CPU compute path exists.
"""
from ._ffi import (RMX_LINSOLVE_LU, RMX_LINSOLVE_PCG, RMX_SCHEME_BDF1, RMX_SCHEME_BDF2, RMX_ST_DIVERGED,
                   RMX_ST_CHART, RMX_ST_LSFAIL, RMX_ST_MAXITER, RMX_ST_NAN, RmxError)
from .scene import (Body, BodyCuboid, ForceCable, ForceGroundCuboid, ForcePointPoint, ForceSpringDamper, Joint, JointFixed, JointFree2D, JointPlanar, JointPrismatic,
                    JointRevolute, JointSpherical, JointFree3D, JointTranslational, JointUniversal, Scene, TaskBDF1PointPos, TaskBDF2PointPos,
                    inertiaCuboid)
from .drivers import (driverRedMaxAdjointBDF1, driverRedMaxAdjointBDF2, driverRedMaxBDF1, driverRedMaxBDF2, plotEnergies,
                      simLoop, taskObjective)
from .scenes import BDF1, BDF2, chain_scene, hand_scene, scenesRedMax, synthetic_inputs, tree_scene

__all__ = [
    'Scene', 'Body', 'BodyCuboid', 'Joint', 'JointRevolute', 'JointFixed', 'JointPrismatic', 'JointPlanar',
    'JointTranslational', 'JointFree2D', 'JointUniversal', 'JointSpherical', 'JointFree3D', 'ForceGroundCuboid', 'ForcePointPoint', 'ForceSpringDamper', 'ForceCable',
    'TaskBDF1PointPos', 'TaskBDF2PointPos', 'inertiaCuboid', 'scenesRedMax', 'chain_scene', 'hand_scene', 'tree_scene',
    'synthetic_inputs', 'BDF1', 'BDF2', 'RmxError', 'driverRedMaxBDF1', 'driverRedMaxBDF2', 'driverRedMaxAdjointBDF1',
    'driverRedMaxAdjointBDF2', 'taskObjective', 'simLoop', 'plotEnergies',
]
