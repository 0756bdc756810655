"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference's detection loss (and, for config C5, its model and
optimiser step) on the synthetic scene of groomed_nms_b200.synthetic.c5_scene, either stock or with
`groomed_nms_b200.install()` active, and records what the GrooMeD branch saw and produced.

    python -m oracle.ref_harness --arm stock|installed --out run.npz [--seed 0] [--batch 2] [--feat 24x80]
                                 [--set key=value ...] [--fake-cuda] [--model]

The reference code is imported from the tree oracle/ref_shim.py finds (baseline/_ref on the GPU box) and is not edited:
the recording is done by two spies bound over the names `differentiable_nms` and `APLoss` in the namespace of the
reference's lib.loss.rpn_3d (lib/loss/rpn_3d.py:13-14); they read the caller's local variables (the per-image
tensors of RPN_3D_loss.forward at :791 and :1117-1131) and delegate to whatever implementation is bound there --
the reference's own in the stock arm, groomed_nms_b200's in the installed arm.  One arm per process (the reference binds
`from lib.groomed_nms import differentiable_nms` at import time).

What is compared by tests/test_gpu_reference_dropin.py: loss, per-loss stats, per image the scores and overlaps that
entered NMS, the rescored scores, the keep lists, scores_after_nms / targets_after_nms, and the gradients of the loss
wrt every network output.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _np(t):
    import torch
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


class Recorder(object):
    def __init__(self):
        self.nms = []           # one dict per image that reached differentiable_nms
        self.ap = []            # one dict per APLoss call
        self.full = {}          # scores_after_nms / targets_after_nms / bbox_weights of the whole batch

    def wrap_nms(self, fn):
        rec = self

        def differentiable_nms(*args, **kw):
            fr = sys._getframe(1).f_locals
            out = fn(*args, **kw)
            img = int(fr["img_index"])
            d = dict(img_index=img, scores_in=_np(kw["scores_unsorted"]), iou_in=_np(kw["iou_unsorted"]),
                     valid=_np(out[0]).astype(np.int64), invalid=_np(out[1]).astype(np.int64), prob=_np(out[2]),
                     fg_inds=_np(fr["fg_inds_tensor"]).astype(np.int64),
                     fg_index_for_nms=_np(fr["fg_index_for_nms"]).astype(np.int64),
                     scores_to_nms_fg=_np(fr["scores_to_nms_img"][fr["fg_inds_tensor"]]),
                     coords_2d_fg=_np(fr["coords_2d_512_img"][fr["fg_inds_tensor"]]),
                     boxes7_fg=np.stack([_np(fr[k][img][fr["fg_inds_tensor"]]) for k in
                                         ("bbox_x3d_raw", "bbox_y3d_raw", "bbox_z3d_raw", "bbox_w3d_raw", "bbox_h3d_raw",
                                          "bbox_l3d_raw", "bbox_ry3d_raw")], 1),
                     gts_val=np.asarray(fr["gts_val"], dtype=np.float64), gts_3d=np.asarray(fr["gts_3d"], dtype=np.float64),
                     corners_b1=_np(fr["corners_3d_b1"]))
            rec.nms.append(d)
            return out
        return differentiable_nms

    def wrap_aploss(self, cls):
        rec = self

        class APLossSpy(object):
            def __init__(self, *a, **k):
                self.inner = cls(*a, **k)

            def __call__(self, logits, targets):
                out = self.inner(logits, targets)
                rec.ap.append(dict(logits=_np(logits), targets=_np(targets), loss=_np(out).reshape(-1)))
                return out
        return APLossSpy


def call_and_capture(rec, forward_code, fn, *args):
    """Run fn(*args); when the frame executing `forward_code` (RPN_3D_loss.forward) returns, keep its batch-wide
    scores_after_nms / targets_after_nms / bbox_weights (lib/loss/rpn_3d.py:297-298,512) in rec.full."""
    def prof(frame, event, arg):
        if event == "return" and frame.f_code is forward_code:
            fr = frame.f_locals
            if "scores_after_nms" in fr:
                B = fr["batch_size"]
                rec.full = dict(scores_after_nms=_np(fr["scores_after_nms"]).reshape(B, -1),
                                targets_after_nms=_np(fr["targets_after_nms"]).reshape(B, -1),
                                bbox_weights=np.asarray(_np(fr["bbox_weights"]), dtype=np.float32).reshape(B, -1))
    sys.setprofile(prof)
    try:
        return fn(*args)
    finally:
        sys.setprofile(None)


def _edict(d):
    from easydict import EasyDict
    e = EasyDict()
    for k, v in d.items():
        e[k] = v
    return e


def build_conf(scene, overrides=None, config="groumd_nms"):
    """The reference's own config module (scripts/config/<config>.py) + the synthetic anchors / normalisation stats."""
    mod = importlib.import_module("scripts.config." + config)
    conf = mod.Config()
    conf.anchors = scene["anchors"]
    conf.bbox_means = scene["bbox_means"]
    conf.bbox_stds = scene["bbox_stds"]
    for k, v in (overrides or {}).items():
        conf[k] = v
    return conf


def build_imobjs(scene):
    out = []
    for gts in scene["gts"]:
        im = _edict(dict(p2=np.array(scene["p2"], dtype=np.float64), scale_factor=float(scene["scale_factor"]),
                         gts=[_edict(dict(cls=g["cls"], ign=bool(g["ign"]), visibility=float(g["visibility"]),
                                          bbox_full=np.array(g["bbox_full"], dtype=np.float64), bbox_3d=list(g["bbox_3d"])))
                              for g in gts]))
        out.append(im)
    return out


def setup(arm, fake_cuda=False):
    """Shim + (installed arm) aliasing, then import of the reference's loss module.  Returns (torch, loss module, recorder)."""
    from oracle import ref_shim
    ref_shim.install()
    import torch
    if fake_cuda:
        ref_shim.fake_cuda()
    else:
        torch.set_default_tensor_type('torch.cuda.FloatTensor')          # what the reference's init_torch does (lib/core.py:899)
    if arm == "installed":
        import groomed_nms_b200
        groomed_nms_b200.install()
    elif arm != "stock":
        raise ValueError(arm)
    L = importlib.import_module("lib.loss.rpn_3d")
    rec = Recorder()
    L.differentiable_nms = rec.wrap_nms(L.differentiable_nms)
    L.APLoss = rec.wrap_aploss(L.APLoss)
    return torch, L, rec


def network_outputs(torch, scene, device):
    """Leaf tensors standing for the network's raw outputs + the derived tensors the loss takes (as the model's forward
    builds them, models/densenet121_3d_dilate_decomp_alpha.py:126-250).  Leaves: cls, bbox_2d_raw, bbox_3d_raw, acc_logit."""
    def leaf(a):
        return torch.from_numpy(np.ascontiguousarray(a)).float().to(device).requires_grad_(True)
    lv = dict(cls=leaf(scene["cls"]), bbox_2d=leaf(scene["bbox_2d"]), bbox_3d=leaf(scene["bbox_3d"]), acc_logit=leaf(scene["acc_logit"]))
    lv = dict(lv)
    B = scene["cls"].shape[0]
    prob = torch.softmax(lv["cls"], dim=2)
    bbox_2d = lv["bbox_2d"] * 1.0                    # non-leaf: bbox_transform_inv scales its deltas in place (lib/rpn_util.py:902-912)
    bbox_3d = lv["bbox_3d"] * 1.0
    acc = torch.sigmoid(lv["acc_logit"]).unsqueeze(2)
    acc.retain_grad()                                # d loss / d acceptance probability (the scores that enter NMS)
    lv["_acc_prob"] = acc
    rois = torch.from_numpy(scene["rois"]).float().to(device)
    anchors = torch.from_numpy(scene["anchors"]).float().to(device)
    rois_3d = anchors[rois[:, 4].long()]
    w = rois[:, 2] - rois[:, 0] + 1.0
    h = rois[:, 3] - rois[:, 1] + 1.0
    cen = torch.stack([rois[:, 0] + 0.5 * w, rois[:, 1] + 0.5 * h], 1)
    rep = lambda t: t.clone().unsqueeze(0).repeat(B, 1, 1)
    return lv, (lv["cls"], prob, bbox_2d, bbox_3d, rep(rois), rep(rois_3d), rep(cen), acc)


def run_loss(arm, scene, overrides=None, fake_cuda=False, config="groumd_nms"):
    torch, L, rec = setup(arm, fake_cuda)
    device = "cpu" if fake_cuda else "cuda"
    conf = build_conf(scene, overrides, config)
    np.random.seed(conf.rng_seed)
    crit = L.RPN_3D_loss(conf, verbose=True)
    lv, (cls, prob, bbox_2d, bbox_3d, rois, rois_3d, rois_cen, acc) = network_outputs(torch, scene, device)
    loss, stats = call_and_capture(rec, L.RPN_3D_loss.forward.__code__, crit, cls, prob, bbox_2d, bbox_3d, build_imobjs(scene),
                                   list(scene["feat_size"]), rois, rois_3d, rois_cen, acc, None)
    loss.backward()
    out = dict(loss=_np(loss).reshape(-1).astype(np.float64), n_nms=np.array([len(rec.nms)]), n_ap=np.array([len(rec.ap)]))
    for s in stats:
        out["stat_" + s["name"]] = np.asarray(_np(s["val"]), dtype=np.float64).reshape(-1)
    for k, t in lv.items():
        g = _np(t.grad) if t.grad is not None else np.zeros(tuple(t.shape), np.float32)
        out["grad_" + k.lstrip("_")] = g[..., 0] if k == "_acc_prob" else g
    for i, d in enumerate(rec.nms):
        for k, v in d.items():
            out["nms%d_%s" % (i, k)] = np.asarray(v)
    for i, d in enumerate(rec.ap):
        for k, v in d.items():
            out["ap%d_%s" % (i, k)] = np.asarray(v)
    for k, v in rec.full.items():
        out["full_" + k] = v
    return out


def run_train_steps(arm, scene, iters=3, overrides=None, fake_cuda=False, config="groumd_nms", seed=0):
    """Config C5 (BASELINE.json configs[4]): the loop body of the reference's scripts/train_rpn_3d.py:131-142 -- its own
    densenet121_3d_dilate_decomp_alpha RPN (random init: `build(conf, 'test')` then .train()), its own
    RPN_3D_loss and its own loss_backprop (backward, clip_grad_value_, SGD step; lib/core.py:99-113) -- on synthetic
    384 x 1280 images and the scene's ground truths, stock or with groomed_nms_b200.install() active.
    Returns the loss of every iteration, gradient probes of the first iteration and the time per iteration."""
    import time
    torch, L, rec = setup(arm, fake_cuda)
    device = "cpu" if fake_cuda else "cuda"
    conf = build_conf(scene, overrides, config)
    np.random.seed(conf.rng_seed)
    torch.manual_seed(seed)
    if not fake_cuda:
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
        # true fp32 in both arms: with TF32 convolutions (torch's default) a 1e-7 difference in the gradient that enters the
        # backbone is re-rounded to 10 mantissa bits layer after layer and the two arms drift apart by 1e-4 of the largest entry
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    model_mod = importlib.import_module("models." + conf.model)
    core = importlib.import_module("lib.core")
    net = model_mod.build(conf, "test")                   # 'test' = no download of pretrained weights; then train mode as the script does
    net.train()
    net.phase = "train"
    if not fake_cuda:
        net = net.cuda()
    opt = torch.optim.SGD(net.parameters(), lr=conf.lr, momentum=conf.momentum, weight_decay=conf.weight_decay)   # lib/core.py:71-77
    crit = L.RPN_3D_loss(conf, verbose=False)
    H, W = scene["feat_size"]
    B = len(scene["gts"])
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    images = torch.randn((B, 3, H * conf.feat_stride, W * conf.feat_stride), generator=g, device="cpu").to(device)
    imobjs = build_imobjs(scene)
    out = {"losses": [], "iter_ms": [], "loss_ms": []}
    probes = {}
    for it in range(iters):
        if not fake_cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        cls, prob, bbox_2d, bbox_3d, feat_size, rois, rois_3d, rois_3d_cen, acc, acc_cls = net(images)          # train_rpn_3d.py:137
        if not fake_cuda:
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        loss, stats = crit(cls, prob, bbox_2d, bbox_3d, imobjs, feat_size, rois, rois_3d, rois_3d_cen, acc, acc_cls)   # :140
        if not fake_cuda:
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        if it == 0:
            loss.backward(retain_graph=True)
            named = dict(net.named_parameters())
            for name in ("acceptance_prob.layer_0.weight", "cls.weight", "bbox_z3d.weight", "prop_feats.0.weight", "base.conv0.weight"):
                probes["grad_" + name] = _np(named[name].grad).reshape(-1)[:4096].copy()
            probes["stat_names"] = np.array([s_["name"] for s_ in stats])
            probes["stat_vals"] = np.array([float(_np(s_["val"]).reshape(-1)[0]) for s_ in stats])
            opt.zero_grad()
        core.loss_backprop(loss, net, opt, conf=conf, iteration=it)                                               # :142
        if not fake_cuda:
            torch.cuda.synchronize()
        t3 = time.perf_counter()
        out["losses"].append(float(loss.detach()))
        out["iter_ms"].append(1e3 * (t3 - t0))
        out["loss_ms"].append(1e3 * (t2 - t1))
    res = dict(losses=np.array(out["losses"]), iter_ms=np.array(out["iter_ms"]), loss_ms=np.array(out["loss_ms"]),
               n_nms=np.array([len(rec.nms)]), nms_sizes=np.array([len(d["scores_in"]) for d in rec.nms]))
    res.update(probes)
    return res


def run_detect(arm, scene, overrides=None, fake_cuda=False, config="groumd_nms"):
    """The inference call site: the reference's im_detect_3d (lib/rpn_util.py:1052-1357) on
    one image, driven by a stand-in `net` that returns the scene's network outputs, stock or with install() active.
    Records the detections as they enter the NMS block (:1258-1290: sorted by score, truncated to nms_topN_pre) and the
    rows im_detect_3d returns.  The reference's gpu_nms is a Cython/CUDA extension for sm_35 that is not rebuilt here; in
    the stock arm its place is taken by the reference's own pure-python twin, lib/nms/py_cpu_nms.py (same +1 convention
    and `>` test, lib/nms/nms_kernel.cu:27-30,71)."""
    import types
    from oracle import ref_shim
    ref_shim.install()
    import torch
    if fake_cuda:
        ref_shim.fake_cuda()
    else:
        torch.set_default_tensor_type('torch.cuda.FloatTensor')
    if arm == "installed":
        import groomed_nms_b200
        groomed_nms_b200.install()
    else:
        py = importlib.import_module("lib.nms.py_cpu_nms")
        sys.modules["lib.nms.gpu_nms"].gpu_nms = lambda dets, thresh, device_id=0: py.py_cpu_nms(dets, thresh)
    R = importlib.import_module("lib.rpn_util")
    device = "cpu" if fake_cuda else "cuda"
    conf = build_conf(scene, overrides, config)
    lv, (cls, prob, bbox_2d, bbox_3d, rois, rois_3d, rois_cen, acc) = network_outputs(torch, scene, device)
    captured = {}

    def spy(fn, name):
        def wrapped(*a, **k):
            fr = sys._getframe(1).f_locals
            if "aboxes" in fr and "pre" not in captured:
                captured["pre"] = dict(aboxes=np.array(fr["aboxes"]), coords_3d=np.array(fr["coords_3d"]), coords_3d_raw=np.array(fr["coords_3d_raw"]),
                                       cls_pred=np.array(fr["cls_pred"]), tracker=np.array(fr["tracker"]))
            out = fn(*a, **k)
            captured["nms_fn"] = name
            captured["keep"] = np.asarray(out[0].numpy() if name == "differentiable_nms" else out, dtype=np.int64)
            return out
        return wrapped
    R.differentiable_nms = spy(R.differentiable_nms, "differentiable_nms")
    R.gpu_nms = spy(R.gpu_nms, "gpu_nms")

    def net(im):                                   # eval-mode outputs of the model (models/...alpha.py:250)
        with torch.no_grad():
            return (cls[:1].detach(), prob[:1].detach(), bbox_2d[:1].detach().clone(), bbox_3d[:1].detach().clone(), list(scene["feat_size"]),
                    rois[0].detach(), acc[:1].detach(), None)
    H, W = scene["feat_size"]
    im = np.zeros((H * 16, W * 16, 3), dtype=np.float32)
    preprocess = lambda x: np.ascontiguousarray(x.transpose(2, 0, 1))
    aboxes = R.im_detect_3d(im, net, conf, preprocess, np.array(scene["p2"], dtype=np.float64))
    out = dict(aboxes_out=np.asarray(aboxes), keep=captured["keep"], nms_fn=np.frombuffer(captured["nms_fn"].encode(), dtype=np.uint8))
    for k, v in captured["pre"].items():
        out["pre_" + k] = v
    return out, conf


def parse_overrides(items):
    ov = {}
    for it in items or []:
        k, v = it.split("=", 1)
        try:
            ov[k] = json.loads(v)
        except ValueError:
            ov[k] = v
    return ov


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", required=True, choices=["stock", "installed"])
    ap.add_argument("--out", required=True)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--feat", default="24x80")
    ap.add_argument("--near-iou", type=float, default=0.3)
    ap.add_argument("--set", action="append", default=[])
    ap.add_argument("--fake-cuda", action="store_true")
    ap.add_argument("--model", action="store_true", help="config C5: the reference's model + loss + optimiser step instead of the loss alone")
    ap.add_argument("--detect", action="store_true", help="the inference call site: the reference's im_detect_3d on one image of the scene")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args(argv)
    from groomed_nms_b200 import synthetic
    H, W = [int(x) for x in a.feat.split("x")]
    scene = synthetic.c5_scene(seed=a.seed, batch=a.batch, feat_size=(H, W), near_iou=a.near_iou)
    if a.detect:
        out, _ = run_detect(a.arm, scene, parse_overrides(a.set), fake_cuda=a.fake_cuda)
        np.savez_compressed(a.out, **out)
        print("ref_harness %s detect: %d detections into NMS (%s), %d kept -> %s" % (a.arm, len(out["pre_aboxes"]), bytes(out["nms_fn"]).decode(),
                                                                                  len(out["aboxes_out"]), a.out))
        return
    if a.model:
        out = run_train_steps(a.arm, scene, iters=a.iters, overrides=parse_overrides(a.set), fake_cuda=a.fake_cuda)
        np.savez_compressed(a.out, **out)
        print("ref_harness %s C5: losses %s, ms per iteration %s (loss alone %s), NMS sizes %s -> %s" % (
            a.arm, np.round(out["losses"], 6).tolist(), np.round(out["iter_ms"], 1).tolist(), np.round(out["loss_ms"], 1).tolist(),
            out["nms_sizes"].tolist(), a.out))
        return
    out = run_loss(a.arm, scene, parse_overrides(a.set), fake_cuda=a.fake_cuda)
    np.savez_compressed(a.out, **out)
    print("ref_harness %s: loss %.6f, %d images through NMS, %d AP-loss calls -> %s" % (a.arm, float(out["loss"][0]), int(out["n_nms"][0]),
                                                                                      int(out["n_ap"][0]), a.out))


if __name__ == "__main__":
    main()
