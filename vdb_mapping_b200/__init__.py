"""vdb_mapping_b200 — B200-native (sm_100a) scan-integration hot path of vdb_mapping.

Only what the path needs: csrc/ (CUDA kernels + C-ABI, built into libvdbm_b200.so), the ctypes
loader, a host-side mirror of the reference's OccupancyVDBMapping interface, synthetic scan
generators and the multi-GPU orchestration. There is no CPU fallback: importing `mapping`
without the built CUDA library raises.
"""
__version__ = "0.1.0"
