"""Host-side mirror of the reference's `lib` package for the GrooMeD-NMS hot path: same module names, function
names, argument meaning and error behaviour as abhi1kumar/groomed_nms `lib/` (groomed_nms, core overlaps, math_3d
corners, nms, nms_others, loss.aploss), implemented on the sm_100a C-ABI.  `groomed_nms_b200.install()` aliases
these modules into `sys.modules['lib.*']` so reference scripts import them unchanged."""
