"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the reference's cross-modal inference / OOD-scoring hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package, and only as the checker or
the CPU baseline -- never as the product path.  The product
(``crossmodal-imu-video-ood-har_b200/``) never imports it.
"""
