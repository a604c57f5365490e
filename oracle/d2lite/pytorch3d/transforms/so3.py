def so3_relative_angle(*a, **k):
    raise NotImplementedError("d2lite: so3_relative_angle is only on the non-disentangled path (roi_heads.py:631)")
