"""multi_orb_slam_b200 — B200-native ORB front end (drop-in for Multi_ORB_SLAM's ORBextractor /
ORBmatcher hot path).  The compute path is the C-ABI CUDA library built from csrc/; there is no
CPU fallback: importing the extractor/matcher modules without the built library raises."""
__version__ = "0.1.0"
