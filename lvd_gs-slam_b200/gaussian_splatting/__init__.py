"""Minimal re-creation of the reference's missing `gaussian_splatting` package: only the renderer shim the
SLAM code imports (`from gaussian_splatting.gaussian_renderer import render`, utils/slam_frontend.py:28,
utils/slam_backend.py:10, utils/init_pose.py:20, utils/eval_utils_0806.py:26).  SURVEY.md section 8(b)."""
