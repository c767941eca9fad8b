"""oracle -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's fused-voxel hot path (GSFusion BiFuser_N, the dense
3D-conv decoder/head, the volume-render regulariser) used as the parity checker.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (co-occ_b200/) never does.
"""
