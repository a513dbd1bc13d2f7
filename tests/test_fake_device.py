"""How far tests/fake_device.py (the oracle behind the slice of the C-ABI that PtzSlam, PTZCamera and the relocaliser call) can be
trusted: GPU tests that are green on a B200 - the per-frame loop against the reference golden, the resident-state ray bookkeeping
against the host path, the six-frame EKF golden, ray add / remove through the device camera, the bundle-adjustment solve against
the reference's own least_squares call - must pass unchanged when
`_lib.get_context` is routed to it.  The fake then stands in for the device in tests/test_zx_cfg1_end_to_end.py, whose own GPU
tests were written after the round's last GPU run."""
import fake_device


def test_hardware_validated_device_tests_pass_on_the_fake(monkeypatch):
    ctx = fake_device.install(monkeypatch)
    import test_tracking_loop as T
    import test_gpu_ekf as E
    import test_ray_bookkeeping as RB
    for c in range(int(T.G["n_runs"])):
        T.test_tracking_loop_device_golden(c)
    E.test_ekf_update_six_frames_golden()
    E.test_resident_state_ray_bookkeeping_matches_host_path()
    for c in range(RB.N_CASES):
        RB.test_add_remove_rays_golden_device_camera(c)
    import test_gpu_ba as B
    B.test_solve_matches_reference_least_squares()           # the reference's own least_squares call: same status, nfev, parameters
    B.test_bundle_adjustment_core_matches_reference()
    B.test_map_add_keyframe_with_ba()
    calls = ctx.lib.calls
    assert calls["update_only"] > 30 and calls["remove_rays"] > 10 and calls["add_rays"] > 10
