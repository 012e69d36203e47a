"""CPU-side checks of the drop-in boundary: the transformer/ mirror has the reference's classes,
state-dict schema and seed-for-seed init; the C-ABI library loads and exports every symbol the
header declares; and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as G
    G.build()
    return G


def test_header_symbols_are_exported(built):
    from tts_b200 import _native
    header = open(os.path.join(ROOT, "include", "tts_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tts_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 14
    lib = ctypes.CDLL(built.LIB)
    for name in declared:
        assert hasattr(lib, name), "libtts_b200.so does not export %s" % name
    assert sorted(_native.exported_symbols()) == declared  # the ctypes binding covers the whole header
    assert _native.load(check_device=False).tts_abi_version() == _native.ABI_VERSION


def test_library_is_sm100a_with_packed_fp32_and_no_legacy_tensor_ops(built):
    sass = subprocess.run(["cuobjdump", "-sass", built.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "FFMA2" in sass               # packed fp32 FMA, the Blackwell fp32 peak path
    assert "HGMMA" not in sass           # no Hopper-only wgmma


def test_struct_layouts_match_the_header(built):
    """sizeof() of the ctypes mirrors equals what the C compiler lays out."""
    from tts_b200 import _native
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "tts_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu", ' \
          'sizeof(TtsGemmEpilogue), sizeof(TtsDecLayerWeights), sizeof(TtsDecoderWeights), sizeof(TtsDecodeState), ' \
          'sizeof(TtsGemmBf16), sizeof(TtsAttnTrain), sizeof(TtsGriffinLim), offsetof(TtsGriffinLim, mag), ' \
          'offsetof(TtsAttnTrain, keep_mask), offsetof(TtsGemmBf16, out_row_offset));return 0;}'
    exe = os.path.join(ROOT, "few-shot-transformer-tts_b200", "build", "sizes")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True,
                   check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    mine = [ctypes.sizeof(c) for c in (_native.GemmEpilogue, _native.DecLayerWeights, _native.DecoderWeights,
                                       _native.DecodeState, _native.GemmBf16, _native.AttnTrain, _native.GriffinLim)]
    mine += [_native.GriffinLim.mag.offset, _native.AttnTrain.keep_mask.offset, _native.GemmBf16.out_row_offset.offset]
    assert sizes == mine


def test_state_dict_schema_and_seeded_init_match_the_reference(built, golden_dir):
    from tts_b200.config import default_hparams
    from transformer import tacotron
    ref = json.load(open(os.path.join(golden_dir, "ref_init_seed0.json")))
    torch.manual_seed(0)                       # train.py:33
    m = tacotron.Tacotron(default_hparams())   # train.py:118
    tacotron.initialize_variables(m)           # train.py:119
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref["tensors"].keys())          # same names, same order
    assert sum(p.numel() for p in m.parameters()) == ref["n_params"]
    for k, v in sd.items():
        want = ref["tensors"][k]
        assert list(v.shape) == want["shape"], k
        np.testing.assert_allclose(float(v.double().sum()), want["sum"], rtol=1e-9, atol=1e-9, err_msg=k)
        np.testing.assert_allclose(float(v.double().abs().sum()), want["abs"], rtol=1e-9, atol=1e-9, err_msg=k)
        np.testing.assert_allclose([float(x) for x in v.flatten()[:3].double()], want["head"], rtol=0, atol=0)
    for step, want in ref["lr_factor"].items():
        assert tacotron.learning_rate_schedule(int(step), default_hparams()) == want


def test_strict_load_of_oracle_weights_and_l2_selection(built, tiny_params):
    from oracle import tts_oracle as O
    from tts_b200.config import hparams_from
    from transformer import tacotron
    cfg, params = tiny_params
    m = tacotron.Tacotron(hparams_from(cfg))
    m.load_state_dict(params, strict=True)      # utils/checkpoint.py:41-44 loads strictly
    batch = O.synth_batch(cfg, batch=3, text_len=20, n_frames=30, seed=6, ragged=True)
    # the L2 term covers exactly the tensors the reference selects by name (tacotron.py:144-146)
    mine = {n for n, p in m.named_parameters() if any(p is q for q in tacotron._l2_selected(m))}
    assert mine == set(O.l2_names(params)) and len(mine) == len(tacotron._l2_selected(m))
    with torch.no_grad():
        outs = O.tacotron_forward(params, cfg, batch)
    # compute_loss is a CUDA kernel path now (values are checked against the oracle in tests/test_gpu_train.py):
    # on CPU tensors it refuses instead of computing somewhere else
    with pytest.raises(RuntimeError, match="no CPU"):
        tacotron.compute_loss(m, batch["mel_targets"], batch["target_lengths"], outs, hparams_from(cfg))


def test_no_cpu_fallback(built, tiny_params):
    """On a CPU model every forward raises instead of silently computing somewhere else."""
    from oracle import tts_oracle as O
    from tts_b200.config import hparams_from
    from transformer import tacotron
    cfg, params = tiny_params
    m = tacotron.Tacotron(hparams_from(cfg)).eval()
    m.load_state_dict(params)
    batch = O.synth_batch(cfg, batch=2, text_len=8, n_frames=6, seed=1)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU"):
        m(**{k: v for k, v in batch.items() if k != "names"})
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU"):
        m.postnet(batch["mel_targets"], batch["target_lengths"])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "few-shot-transformer-tts_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), os.path.join(base, f)


def test_packed_rows_layout():
    """Host packing of the pipelined kernel's weight operand (tts_b200.h pk_*): k-permutation per 16-float chunk,
    K-split-major slices, constants in the row tail."""
    import torch
    from tts_b200.engine import pack_rows
    g = torch.Generator().manual_seed(3)
    w = torch.randn(10, 64, generator=g)
    c, s = torch.randn(10, generator=g), torch.randn(10, generator=g)
    p = pack_rows(w, c, s)
    assert p.shape == (1, 10, 80)
    for n in (0, 7):
        for chunk in range(4):
            for t in range(4):
                for i in range(4):
                    assert p[0, n, chunk * 16 + 4 * t + i] == w[n, chunk * 16 + t + 4 * i]
    assert torch.equal(p[0, :, 64], c) and torch.equal(p[0, :, 65], s) and float(p[0, :, 66:].abs().max()) == 0.0
    q = pack_rows(w, ksplit=2)
    assert q.shape == (2, 10, 48)
    assert q[1, 3, 4 * 2 + 1] == w[3, 32 + 2 + 4 * 1] and q[0, 3, 16 + 4 * 3 + 0] == w[3, 16 + 3]
    assert float(q[:, :, 32:].abs().max()) == 0.0


def test_unmodified_reference_synthesize_runs_against_the_mirror(built, tiny_params):
    """The reference's own synthesize.py (unmodified, imported from /root/reference) driving THIS transformer/ package:
    import wiring, the hp / batch-dict plumbing and the call sequence up to the first kernel call, which on this
    CPU-only container must fail loudly with our "no CPU path" error (there is no fallback to hide behind).  The GPU run
    recipe is in INTEGRATION.md; the same loop on CUDA is tests/test_gpu_parity.py::test_module_api_unchanged_synthesis_loop."""
    import importlib
    import sys
    from unittest import mock
    ref = os.environ.get("TTS_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref, "synthesize.py")):
        pytest.skip("reference checkout not present (GPU box)")
    from oracle import tts_oracle as O
    from transformer import tacotron
    cfg, params = tiny_params
    saved = {k: sys.modules.get(k) for k in ("synthesize", "hyperparams", "utils", "utils.infolog", "utils.audio", "librosa", "soundfile",
                                             "fastdtw", "matplotlib", "matplotlib.pyplot", "editdistance")}
    for name in ("librosa", "librosa.filters", "librosa.effects", "soundfile", "fastdtw", "matplotlib", "matplotlib.pyplot", "editdistance"):
        sys.modules.setdefault(name, mock.MagicMock())
    sys.path.append(ref)      # AFTER the repo: `transformer` resolves to this package, everything else to the reference
    try:
        synth = importlib.import_module("synthesize")
        assert os.path.samefile(os.path.dirname(synth.__file__), ref)
        import transformer
        assert "few-shot-transformer-tts_b200" in transformer.__file__
        hp = synth.hp
        for k, v in vars(cfg).items():
            hp.set_hparam(k, v)
        m = tacotron.Tacotron(hp)                                  # reference HParams object -> our constructors
        m.load_state_dict(params, strict=True)
        m.eval()
        batch = O.synth_batch(cfg, batch=2, text_len=12, n_frames=4, seed=1)
        data = {k: batch[k] for k in ("inputs", "input_lengths", "input_spk_ids", "input_language_vecs", "names")}
        with pytest.raises(RuntimeError, match="no CPU"):
            synth.eval_batch(m, data, use_bar=False, bar_interval=-1)
    finally:
        sys.path.remove(ref)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            if getattr(sys.modules[k], "__file__", "") and ref in str(getattr(sys.modules[k], "__file__", "")):
                sys.modules.pop(k, None)
