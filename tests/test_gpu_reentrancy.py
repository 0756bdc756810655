"""The C-ABI keeps no mutable process-wide state (include/groomed_nms_b200.h): tuning knobs travel per call in a
gnms_launch_opts.  Two host threads drive two streams with DIFFERENT options at the same time -- one the register-direct
matrix kernel with CTAs that retire after a tile, the other the TMA kernel with persistent CTAs, plus full forwards with
different rank / election methods -- and every result must equal the single-threaded one bit for bit."""
import threading

import numpy as np
import pytest
import torch

from gpu_util import cuda

pytestmark = pytest.mark.gpu


def test_two_threads_two_streams_different_opts():
    from groomed_nms_b200 import _lib, ops, synthetic
    lib = _lib.load()
    B, N = 3, 1536
    b7 = np.stack([synthetic.config_c3(seed=50 + i, n=N, k=12)[0] for i in range(B)])
    sc = np.stack([synthetic.config_c3(seed=50 + i, n=N, k=12)[1] for i in range(B)])
    rec = ops.box3d_records(ops.corners_from_boxes7(cuda(b7).view(B * N, 7))).view(B, N, 8).contiguous()
    scores = cuda(sc)
    p = ops.make_params(group_size=30)
    want_ov = torch.empty((B, N, N), device="cuda")
    _lib.check(lib.gnms_overlap3d_batched_f32(ops._p(rec), N, B, ops._p(want_ov), 1, 1, None), "ref overlap")
    want = ops.forward_boxes(scores, rec, _lib.BOX_3D_REC, p, generalized=True, affine=True)
    torch.cuda.synchronize()
    variants = [dict(mat=_lib.launch_opts(matrix_kernel=_lib.MATRIX_KERNEL_DIRECT, tiles_per_cta=4),
                     fwd=_lib.launch_opts(rank_method=_lib.RANK_COUNT, election=_lib.ELECT_DIRECT)),
                dict(mat=_lib.launch_opts(matrix_kernel=_lib.MATRIX_KERNEL_TMA, tiles_per_cta=0),
                     fwd=_lib.launch_opts(rank_method=_lib.RANK_SORT, election=_lib.ELECT_MASK))]
    errors = []

    def worker(v):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for it in range(25):
                    ov = torch.full((B, N, N), -1.0, device="cuda")
                    _lib.check(lib.gnms_overlap3d_batched_ex_f32(ops._p(rec), N, B, ops._p(ov), 1, 1, _lib.opts_ref(v["mat"]),
                                                                 st.cuda_stream), "overlap")
                    got = ops.forward_boxes(scores, rec, _lib.BOX_3D_REC, p, generalized=True, affine=True, opts=v["fwd"],
                                            private_ws=True)
                    st.synchronize()
                    if not torch.equal(ov, want_ov):
                        errors.append("matrix differs in iteration %d" % it)
                    for f in ("order", "lead", "prob", "counts"):
                        if not torch.equal(getattr(got, f), getattr(want, f)):
                            errors.append("%s differs in iteration %d" % (f, it))
        except Exception as e:                                        # noqa: BLE001 - reported through `errors`
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(v,)) for v in variants]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
