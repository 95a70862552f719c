"""shg_vqa_b200 — B200-native (sm_100a) hot path of SHG-VQA behind the reference's Python module boundary."""
__version__ = "0.1.0"
