from .lidar import (Lidar2D, ClippedLidar2D, RandomPoseLidarDataset, TrajectoryLidarDataset,
                    OnlineTrajectoryLidarDataset, interpolate_waypoints)

__all__ = ["Lidar2D", "ClippedLidar2D", "RandomPoseLidarDataset", "TrajectoryLidarDataset",
           "OnlineTrajectoryLidarDataset", "interpolate_waypoints"]
