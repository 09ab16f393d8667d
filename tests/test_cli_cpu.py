"""CLI host logic that needs no GPU: the teacher-utterance loader of generate_synthesis (reference
src/script/generate_synthesis.py:86-90, src/common/data_utils.py:55-59)."""
import os
import sys

import numpy as np
import pytest

import fac_via_ppg_b200
from fac_via_ppg_b200.script import generate_synthesis


def test_npy_ppg_is_loaded_as_float32(tmp_path):
    path = str(tmp_path / "ppg.npy")
    np.save(path, np.random.rand(7, 5816))
    ppg = generate_synthesis.load_teacher_ppg(path)
    assert ppg.shape == (7, 5816) and ppg.dtype == np.float32
    np.save(path, np.zeros(5))
    with pytest.raises(ValueError):
        generate_synthesis.load_teacher_ppg(path)


def test_recording_without_reference_front_end_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.delenv("FAC_REFERENCE_SRC", raising=False)
    with pytest.raises(SystemExit) as exc:
        generate_synthesis.load_teacher_ppg(str(tmp_path / "teacher.wav"))
    assert "FAC_REFERENCE_SRC" in str(exc.value)


@pytest.fixture
def alias_sandbox():
    """The drop-in aliases live in sys.modules: keep them from leaking into tests that import the real reference."""
    names = lambda: [n for n in sys.modules if n.split(".")[0] in ("common", "waveglow", "ppg")]   # noqa: E731
    saved = {n: sys.modules[n] for n in names()}
    yield
    for n in names():
        del sys.modules[n]
    sys.modules.update(saved)


def test_recording_uses_the_reference_front_end_despite_the_aliases(tmp_path, monkeypatch, alias_sandbox):
    """The drop-in aliases bind `common` to this package (no data_utils there); the reference's own `common` and
    `ppg` packages must still be importable for get_ppg, and the aliases must be back afterwards."""
    src = tmp_path / "src"
    (src / "common").mkdir(parents=True)
    (src / "ppg").mkdir()
    (src / "common" / "__init__.py").write_text("")
    (src / "common" / "utils.py").write_text("def load_filepaths(p):\n    return [p]\n")
    (src / "common" / "data_utils.py").write_text(
        "import numpy as np\nfrom common.utils import load_filepaths\n"
        "def get_ppg(path, deps):\n    assert load_filepaths(path) == [path] and deps.tag == 'deps'\n"
        "    return np.full((3, 5816), 0.5)\n")
    (src / "ppg" / "__init__.py").write_text("class DependenciesPPG:\n    tag = 'deps'\n")
    fac_via_ppg_b200.install_aliases(force=True)
    before = sys.modules["common"]
    monkeypatch.setenv("FAC_REFERENCE_SRC", str(src))
    ppg = generate_synthesis.load_teacher_ppg(str(tmp_path / "teacher.wav"))
    assert ppg.shape == (3, 5816) and ppg.dtype == np.float32 and float(ppg[0, 0]) == 0.5
    assert sys.modules["common"] is before and "ppg" not in sys.modules
    assert str(src) not in sys.path
    from common.utils import get_inference      # the drop-in again  # noqa: F401
    assert os.path.basename(sys.modules["common.utils"].__file__) == "utils.py"
    assert "fac_via_ppg_b200" in sys.modules["common.utils"].__file__


def test_sparse_ppg_host_pruning_round_trip():
    """ops.SparsePPG.from_dense_host: top-k per frame, ascending channel order, zero padding; dense() restores
    exactly the kept entries."""
    import torch
    from fac_via_ppg_b200.ops import SparsePPG
    g = torch.Generator().manual_seed(0)
    dense = torch.zeros(2, 50, 7)
    dense[0, [3, 17, 40], 0] = torch.tensor([0.2, 0.5, 0.3])
    dense[1, :, 2] = torch.softmax(torch.randn(50, generator=g), 0)
    sp = SparsePPG.from_dense_host(dense, k=4)
    assert sp.shape == (2, 7, 4) and sp.n_symbols == 50
    assert sp.indices[0, 0].tolist() == [3, 17, 40, 0] and sp.values[0, 0].tolist()[3] == 0.0
    kept = sp.dense()
    assert torch.equal(kept[0], dense[0])                                   # <= k entries per frame: lossless
    top4 = dense[1, :, 2].topk(4)
    assert sorted(sp.indices[1, 2].tolist()) == sorted(top4.indices.tolist())
    assert sp.indices[1, 2].tolist() == sorted(sp.indices[1, 2].tolist())
    assert abs(float(kept[1, :, 2].sum()) - float(top4.values.sum())) < 1e-6
    thr = SparsePPG.from_dense_host(dense, k=4, threshold=0.25)
    assert thr.values[0, 0].tolist() == pytest.approx([0.5, 0.3, 0.0, 0.0]) and thr.indices[0, 0].tolist() == [17, 40, 0, 0]
    with pytest.raises(ValueError):
        SparsePPG(torch.zeros(1, 2, 65, dtype=torch.int32), torch.zeros(1, 2, 65), 100)
