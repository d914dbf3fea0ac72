"""Host-side mirror of the reference interface: validation and error behaviour that must hold before any
GPU work starts (sklearn _nmf.py:61-76, 1205-1221; 04_align_n_nmf.py:212-213)."""
import numpy as np
import pytest

from exemplars_vc_b200 import nmf, sharding, synth


def _xa(dtype=np.float64):
    rng = np.random.default_rng(0)
    return rng.random((6, 5)).astype(dtype), rng.random((7, 5)).astype(dtype)


def test_reference_configuration_only():
    X, A = _xa()
    with pytest.raises(NotImplementedError):
        nmf.non_negative_factorization(X, H=A, init="custom", update_H=True, solver="mu")
    with pytest.raises(NotImplementedError):
        nmf.non_negative_factorization(X, H=A, init="custom", update_H=False, solver="cd")
    with pytest.raises(ValueError):
        nmf.non_negative_factorization(X, H=A, init=None, update_H=False, solver="mu")
    with pytest.raises(NotImplementedError):
        nmf.non_negative_factorization(X, H=A, init="custom", update_H=False, solver="mu", beta_loss="itakura-saito")
    with pytest.raises(ValueError):
        nmf.non_negative_factorization(X, H=None, init="custom", update_H=False, solver="mu")


def test_dtype_mismatch_is_a_type_error():
    X, A = _xa()
    with pytest.raises(TypeError, match="H should have the same dtype as X"):
        nmf.non_negative_factorization(X, H=A.astype(np.float32), init="custom", update_H=False, solver="mu")


def test_shape_and_sign_errors():
    X, A = _xa()
    with pytest.raises(ValueError, match="wrong second dimension"):
        nmf.non_negative_factorization(X, H=A[:, :4], init="custom", update_H=False, solver="mu")
    with pytest.raises(ValueError, match="wrong first dimension"):
        nmf.non_negative_factorization(X, H=A, n_components=3, init="custom", update_H=False, solver="mu")
    Xn = X.copy(); Xn[0, 0] = -1.0
    with pytest.raises(ValueError, match="Negative values"):
        nmf.non_negative_factorization(Xn, H=A, init="custom", update_H=False, solver="mu")


def test_non_finite_inputs_raise_like_check_array():
    """sklearn's check_array rejects NaN / inf before anything else (a NaN is not < 0, so the sign check alone
    would let it through)."""
    X, A = _xa()
    Xn = X.copy(); Xn[1, 2] = np.nan
    with pytest.raises(ValueError, match="Input X contains NaN"):
        nmf.non_negative_factorization(Xn, H=A, init="custom", update_H=False, solver="mu")
    An = A.copy(); An[0, 0] = np.inf
    with pytest.raises(ValueError, match="Input H contains infinity"):
        nmf.non_negative_factorization(X, H=An, init="custom", update_H=False, solver="mu")


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    X, A = _xa()
    with pytest.raises(RuntimeError, match="no CPU path"):
        nmf.non_negative_factorization(X, H=A, init="custom", update_H=False, solver="mu",
                                       beta_loss="kullback-leibler")


def test_exemplar_range_covers_and_aligns():
    for N, world in [(200000, 8), (20000, 2), (1000, 4), (130, 2), (128, 3)]:
        rs = [sharding.exemplar_range(N, r, world) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == N
        for (a0, a1), (b0, b1) in zip(rs, rs[1:]):
            assert a1 == b0 and a0 <= a1
        for a0, _ in rs:
            assert a0 % 128 == 0 or a0 == N


def test_partition_utterances_is_balanced_and_deterministic():
    lengths = synth.utterance_lengths(synth.BASE_SEED, 256)
    for world in (1, 2, 4, 8):
        plan = sharding.partition_utterances(lengths, world)
        assert sorted(i for p in plan for i in p) == list(range(256))
        loads = [int(sum(lengths[i] for i in p)) for p in plan]
        assert max(loads) - min(loads) <= int(lengths.max())
        assert plan == sharding.partition_utterances(list(lengths), world)


def test_workloads_match_baseline_configs():
    c = synth.CONFIGS
    assert (c["single_utterance_20k"].F, c["single_utterance_20k"].N, c["single_utterance_20k"].T,
            c["single_utterance_20k"].iterations) == (513, 20000, 1000, 500)
    assert (c["large_dictionary_200k"].N, c["large_dictionary_200k"].T) == (200000, 2000)
    assert (c["context_stacked_50k"].F, c["context_stacked_50k"].N) == (2565, 50000)
    L = synth.utterance_lengths(synth.BASE_SEED + 2, 256)
    assert L.min() >= 400 and L.max() <= 600 and len(L) == 256


def test_script_level_signatures():
    import inspect
    from exemplars_vc_b200 import align_n_nmf, align_n_nmf_pytorch, conversion
    from exemplars_vc_b200.nmf_tool.nmf import NMF
    assert str(inspect.signature(align_n_nmf._factorize)) == "(X, W, beta_loss='kullback-leibler', tol=0.0001)"
    assert list(inspect.signature(align_n_nmf.factorize).parameters) == ["tobe_converted", "src_feat"]
    assert list(inspect.signature(align_n_nmf.convert).parameters) == ["H", "tar_feat", "residual"]
    assert list(inspect.signature(align_n_nmf_pytorch.convert).parameters) == ["H", "tar_feat"]
    assert list(inspect.signature(conversion._get_conversion_data).parameters) == ["audiodatum", "fs", "refine_f0"]
    assert list(inspect.signature(conversion.io_load_from_pickle).parameters) == ["speaker"]
    p = inspect.signature(NMF.__init__).parameters
    assert [p[k].default for k in ("max_iter", "learning_rate", "display_step", "optimizer", "initW")] == \
        [200, 0.01, 10, "mu", False]
    assert list(inspect.signature(NMF.fit_transform).parameters)[:5] == ["self", "X", "r_components", "initW", "givenW"]
    with pytest.raises(NotImplementedError):
        NMF().fit_transform(np.ones((3, 2)), 2, False, 0)


def test_work_decomposition_of_both_contractions(tmp_path):
    """tests/host/plan_check.cu (host-only, no GPU): for the BASELINE shapes and 3000 random ones, plan_c1 +
    decode_item give every (row group, frame tile) its whole K range exactly once, no work item is empty, split
    indices stay inside the partial buffer, and the half-width tail items of contraction 2 tile each frame tile."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "plan_check")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                    os.path.join(root, "tests", "host", "plan_check.cu"), "-ldl"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and " 0 failures" in r.stdout, r.stdout[-2000:]
