"""The native plugins EXECUTING under the unmodified reference's own L1-L5 code on the GPU (tier 1 of INTEGRATION.md).

The reference tree comes from /root/reference (build container) or its verbatim install baseline/_ref (GPU box; written
by tools/install_reference.sh from __graft_entry__.build()).  Skipped when neither exists."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_reference_code_runs_the_native_plugins_bit_exact_on_the_gpu():
    """quantizers/quantization_manager.py:41-42,55-90 and quantizers/fake_quantize.py:43-51 (the reference's files) drive
    the CUDA kernels through the registry: YOLOv8n, 57 fused layers, W8A8 and W4A8, calibration -> learnable scales ->
    three SGD steps.  Against the reference's own plugins on the same GPU (same cuDNN algorithms, CUDA-tensor scales):
    calibrated scales / zero-points and LSQ initialisations identical, every fused layer's output, the input gradient and
    every weight / bias gradient bit-identical, every scale gradient within 1e-5 x the sum of |terms| of its reductions
    (north_star's fp32 bar; the reference adds two fp32 full-tensor sums in ATen's order, the kernels' sums are checked
    against the fp64 oracle in test_gpu_kernels.py), losses of the three steps equal to 1e-6 relative."""
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present (neither /root/reference nor baseline/_ref)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_tier1_driver.py")], cwd="/tmp",
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-1500:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("TIER1_JSON ")][-1]
    for r in json.loads(line[len("TIER1_JSON "):]):
        assert r["reference_plugin"] == "quantizers.uniform" and r["native_plugin"] == "vsiquantization_b200.quantizers.uniform", r
        assert r["n_fused"] == 57 and r["layers"] == 57
        assert r["scale_devices"] == ["cuda:0"], r["scale_devices"]
        assert r["calibrated_equal"], "post-calibration scales / zero-points differ from the reference's"
        assert r["init_equal"], "LSQ initialisation differs"
        assert r["layers_with_output_mismatch"] == [], r["layers_with_output_mismatch"]
        assert r["dx_equal"], "input gradient differs"
        assert r["weight_bias_grads_with_mismatch"] == [], r["weight_bias_grads_with_mismatch"]
        assert r["n_scale_grads"] == 114
        assert r["scale_grad_worst_err_over_mass"] <= 1e-5, r["scale_grad_worst_err_over_mass"]
        for la, lb in zip(r["losses_reference"], r["losses_native"]):
            assert abs(la - lb) <= 1e-6 * abs(la), (r["losses_reference"], r["losses_native"])
