from .circular_motion import CircularMotion, CircleEnvAtacom, CircleEnvErrorCorrection
from .collision_avoidance import PointGoalReach, PointReachAtacom
from .air_hockey import JointSpaceEnv, AirHockeyIiwaAtacom, AirHockeyPlanarAtacom

__all__ = ["CircularMotion", "CircleEnvAtacom", "CircleEnvErrorCorrection", "PointGoalReach", "PointReachAtacom",
           "JointSpaceEnv", "AirHockeyIiwaAtacom", "AirHockeyPlanarAtacom"]
