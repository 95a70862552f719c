"""Import the UNMODIFIED reference (read-only at /root/reference) on CPU.

Test infrastructure only.  Used by tools/gen_golden.py (build container, reads
/root/reference) to emit the golden vectors under tests/golden/, and by
tests/test_gpu_training_boundary.py::test_reference_train_loop_on_hot_path, which
runs the reference's unmodified train loop on the GPU box from a copy staged by
tools/stage_reference.py under baseline/_ref/ (git-ignored; the test skips when
neither location exists).  Nothing in the product path imports this module.

Recipe: set sys.argv to the README flags before importing
src.param (param.py:201 parses at import); stub the unused third-party imports
(modeling_capsbert.py:30 timm, file_utils.py:16-18 boto3, vis_utils.py:7-8
ipyplot/matplotlib, agqa_data.py:4-5 ffmpeg/h5py, data_transforms.py:5
pytorchvideo); replace the three network-touching constructors
(modeling_capsbert.py:1653 from_pretrained, tokenization.py:136, video_encoder.py:43).
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference/AGQA"
if not os.path.isdir(os.path.join(REF_ROOT, "src")):
    REF_ROOT = os.path.join(_REPO, "baseline", "_ref", "AGQA")      # staged copy (tools/stage_reference.py)

README_FLAGS = [
    "--llayers", "5", "--xlayers", "2", "--rlayers", "5", "--noCaps",
    "--crossAttnType", "cross", "--batchSize", "32", "--optim", "bert",
    "--lr", "1e-5", "--fromScratch", "--LossHGPerFrame", "--backbone", "slow_r50",
]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def import_reference(task="--taskHGQA"):
    """Returns a namespace with the reference modules. Only one task per process
    (the reference's argparse singleton is evaluated at import)."""
    import torch

    if "src.param" in sys.modules:
        ns = sys.modules["__shg_ref_ns__"]
        assert ns.task == task, "reference already imported with another task flag"
        return ns
    sys.argv = ["ref"] + README_FLAGS + [task]
    sys.path.insert(0, REF_ROOT)
    _stub("timm")
    _stub("boto3")
    _stub("botocore")
    _stub("botocore.exceptions", ClientError=Exception)
    _stub("ipyplot")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("ffmpeg")
    _stub("h5py")
    _stub("pytorchvideo")
    names = ["ApplyTransformToKey", "ShortSideScale", "UniformTemporalSubsample",
             "UniformCropVideo", "Normalize", "AugMix", "RandAugment", "Permute",
             "RandomShortSideScale", "Div255", "create_video_transform"]
    _stub("pytorchvideo.transforms", **{n: object for n in names})

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return object
    for extra in ("pytorchvideo.data", "pytorchvideo.data.encoded_video",
                  "pytorchvideo.transforms.functional", "torchvision.transforms._transforms_video"):
        if extra not in sys.modules:
            sys.modules[extra] = _Any(extra)

    import src.lxrt.modeling_capsbert as mc
    import src.lxrt.tokenization as tok

    def _from_pretrained(cls, name, state_dict=None, cache_dir=None, *inputs, **kwargs):
        return cls(mc.BertConfig(30522), *inputs, **kwargs)
    mc.BertPreTrainedModel.from_pretrained = classmethod(_from_pretrained)
    tok.BertTokenizer.from_pretrained = classmethod(lambda cls, *a, **k: None)

    import src.video_encoder as ve

    class _IdentityBackbone(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def encode(self, x):
            return x
    ve.VideoBackbone = _IdentityBackbone
    import src.lxrt.entry as entry
    import src.lxrt.transformer as transformer
    import src.lxrt.matcher as matcher
    import src.tasks.agqa_model as agqa_model
    agqa_model.VideoBackbone = _IdentityBackbone
    import src.tasks.agqaHGQA as hgqa

    ns = types.SimpleNamespace(task=task, mc=mc, entry=entry, transformer=transformer,
                               matcher=matcher, agqa_model=agqa_model, hgqa=hgqa,
                               args=sys.modules["src.param"].args)
    sys.modules["__shg_ref_ns__"] = ns
    return ns


def build_reference_model(ns, num_answers=171, num_queries=128, num_classes=456, num_actions=157):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = ns.agqa_model.AGQAModel(num_answers, num_queries=num_queries,
                                    num_classes=num_classes, num_actions=num_actions)
    return m.eval()


def reference_loss_object(ns):
    """An AGQA task object with only the fields loss_labels needs (agqaHGQA.py:203-229)."""
    A = object.__new__(ns.hgqa.AGQA)
    A.background_idx = 0
    A.clip_len = 16
    return A
