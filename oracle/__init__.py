"""TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy, fp32) of the reference's GrooMeD-NMS hot path plus the import shim that runs the
real reference in the build container.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package; the product (groomed_nms_b200/) never does.
"""
