"""Parity gate at the BENCHMARKED shape (SURVEY §7.1 "goldens at small and full shapes"; BASELINE north_star:
"identical synthetic 1024x2048 Cityscapes-shaped batches"): one training step of BASELINE configs[1], [2], [3] at
2 (3) x 1024 x 2048 on the benchmarked dense arm (3xTF32 forward / TF32 backward) against the CPU oracle with the
oracle's random draws replayed — losses within 1e-4, RPN labels / sampled anchors bit-exact, sampled ROIs identical
(labels, domains, boxes up to equal-score ties), and, for configs[2] (every head active), gradient probes within the
mixed arm's stated gradient tolerance.  An oracle step costs 10-60 s of host time."""
import json

import pytest

from fullsize_parity import run, verdict

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(1500)
@pytest.mark.parametrize("index,with_grads", [(1, False), (2, True), (3, False)])
def test_full_size_training_step_matches_oracle(index, with_grads):
    rep = run(index, dense="mixed", with_grads=with_grads)
    print(json.dumps(rep, default=str))
    bad = verdict(rep, loss_tol=1e-4)
    assert not bad, bad
    if with_grads:
        # TF32 backward over fp32-grade activations: 1e-2 of the global gradient norm (tests/test_gpu_model.py TOL)
        assert rep["grad_global"] < 1e-2, rep["grad_global"]
        trunk = [v for k, v in rep["grads"].items() if "da_heads" not in k]
        assert max(trunk) < 5e-2, rep["grads"]
