"""The reference's ENTRY POINT, unmodified, on top of the CUDA module: ``train.py`` (upstream train.py:1-154) is imported
from the staged reference files and its ``main()`` is executed -- ``drin/data.py`` loads a synthetic cache directory
written under the reference's own file names, ``MELModel`` / ``TripletLoss`` / ``TopkAccuracy`` / ``EpochLogger`` are the
reference's, Lightning is the 100-line stand-in of ``oracle/lightning_stub.py`` (the package is absent from this image).
Once with the reference's ``drin.model.Model`` on the CPU, once with ``drin_b200.install_as_reference_module()`` on the
GPU (the one-line switch of INTEGRATION.md): the logged losses, top-k accuracies and final weights must agree.
Skipped when the reference files are not staged (``python oracle/make_ref.py``)."""
import io
import os
import re
import sys
from contextlib import redirect_stdout

import pytest
import torch

import drin_b200
from drin_b200.store import synthetic_tables
from oracle import lightning_stub, ref_import
from tests.test_store import write_cache_dir

HAVE = ref_import.available() and os.path.isfile(os.path.join(ref_import.REFERENCE_ROOT, "train.py"))
needs_ref = pytest.mark.skipif(not HAVE, reason="reference train.py / drin/data.py not staged (python oracle/make_ref.py)")


ENTITY_TABLES = ("entity_text_feature", "entity_text_mask", "entity_image_feature", "entity_object_feature",
                 "entity_object_score")
SHAPES = {"wikidiverse": dict(cands=10, le=64, kw={}),
          # WikiMEL layout: one entity table shared by the splits, candidates found through qid2idx (drin/data.py:88-93)
          "wikimel": dict(cands=5, le=16, kw=dict(num_entities=24, entity_tokens=16, mention_tokens=32))}


def _write_splits(root, dataset):
    sh = SHAPES[dataset]
    shared = None
    for split, n, seed in (("train", 20, 1), ("valid", 8, 2), ("test", 8, 3)):       # 20 = 2 x 8 + 4: a short last batch
        t = synthetic_tables(dataset, n, seed, sh["cands"], **sh["kw"])
        if dataset == "wikimel":
            shared = shared or {k: t[k] for k in ENTITY_TABLES}
            t.update(shared)
        write_cache_dir(root, dataset, t, split, sh["cands"] + 1)


def _run_reference_main(root, cuda_model: bool, dataset="wikidiverse"):
    """Import the reference's train.py under a patched common.args and run main(); returns (model, stdout)."""
    sys.modules.pop("train", None)
    sh = SHAPES[dataset]
    ref_import.load(dataset, sh["cands"], sh["le"], preprocess_dir=root, mention_mmap=None, entity_mmap=None,
                    dataloader_workers=0, batch_size=8, shuffle_train_data=False, num_epoch=2, test_epoch_interval=1,
                    use_device="cuda" if cuda_model else "cpu", output_test_result=False, profiling=False, debug=False,
                    test_only=False, seed=0, model_type="drin", metrics_topk=[1, 3, 5], acc_correction=[0, 0, 0])
    lightning_stub.install()
    if cuda_model:
        drin_b200.install_as_reference_module()
    import train                                         # noqa: the reference's entry point, byte for byte
    if cuda_model:
        assert train.model_module.Model is drin_b200.Model
    created = []
    make = train.model_module.Model
    train.model_module.Model = lambda: (created.append(make()) or created[-1])
    out = io.StringIO()
    try:
        with redirect_stdout(out):
            train.main()
    finally:
        train.model_module.Model = make
        for name in [m for m in sys.modules if m.split(".")[0] in ("common", "baselines", "drin", "train")]:
            del sys.modules[name]
    return created[0], out.getvalue()


def _logged(text):
    steps = [s for s in re.split(r"[\r\n]", text) if "loss:" in s]
    losses = [float(re.search(r"loss: ([-0-9.eE]+|nan)", s).group(1)) for s in steps]
    top1 = [float(re.search(r"top-1: ([-0-9.eE]+|nan)", s).group(1)) for s in steps]
    return losses, top1


@needs_ref
def test_stub_runs_the_reference_entry_point_on_cpu(tmp_path):
    root = str(tmp_path) + os.sep
    _write_splits(root, "wikidiverse")
    model, text = _run_reference_main(root, cuda_model=False)
    losses, top1 = _logged(text)
    # 2 blocks x (3 train + 1 valid + 1 test) steps, banners of EpochLogger, the closing line of main()
    assert len(losses) == 10 and all(l == l for l in losses)
    assert text.count("***** Epoch") == 6 and "Training completed" in text
    assert type(model).__module__ == "drin.model"


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("dataset", ["wikidiverse", "wikimel"])
def test_reference_train_py_runs_unmodified_on_the_cuda_module(tmp_path, dataset):
    root = str(tmp_path) + os.sep
    _write_splits(root, dataset)
    ref_model, ref_text = _run_reference_main(root, cuda_model=False, dataset=dataset)
    our_model, our_text = _run_reference_main(root, cuda_model=True, dataset=dataset)
    assert isinstance(our_model, drin_b200.Model) and next(our_model.parameters()).is_cuda
    (l_ref, t_ref), (l_our, t_our) = _logged(ref_text), _logged(our_text)
    assert len(l_ref) == len(l_our) == 10
    assert all(abs(a - b) <= 3e-5 + 2e-4 * abs(b) for a, b in zip(l_our, l_ref)), (l_our, l_ref)    # logged with 5 decimals
    assert t_our == t_ref                                                                          # running top-1 accuracy
    assert our_text.count("***** Epoch") == ref_text.count("***** Epoch") == 6
    # six Adam steps from bit-identical initial weights: entries whose gradient is noise-level may step differently
    for (k, p), (_, r) in zip(our_model.named_parameters(), ref_model.named_parameters()):
        diff = (p.detach().cpu() - r.detach()).abs()
        assert float(diff.mean()) < 2e-5 and float((diff > 5e-4).double().mean()) < 5e-3, k
