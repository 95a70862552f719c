"""Runs the train-mode forward twice and reports which saved tensors differ bitwise (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_dropout import _trainer, _batch

def flat(prefix, obj, acc):
    if torch.is_tensor(obj):
        acc[prefix] = obj
    elif isinstance(obj, (list, tuple)):
        for i, o in enumerate(obj):
            flat("%s.%d" % (prefix, i), o, acc)
    elif isinstance(obj, dict):
        for k, o in obj.items():
            flat("%s.%s" % (prefix, k), o, acc)

inp, batch = _batch()
for drop in (False, True):
    for streams in (True, False):
        res = []
        for rep in range(2):
            tr = _trainer(drop, seed=5)
            tr.engine.use_streams = streams
            tr._advance_seed()
            out, ctx = tr.engine.forward(inp, train=True)
            torch.cuda.synchronize()
            acc = {}
            flat("out", out, acc); flat("ctx", ctx, acc)
            res.append({k: v.clone() for k, v in acc.items()})
        bad = [k for k in res[0] if res[0][k].shape == res[1][k].shape and not torch.equal(res[0][k], res[1][k])]
        print("dropout", drop, "streams", streams, "tensors", len(res[0]), "differing", len(bad))
        for k in bad[:12]:
            a, b = res[0][k].float(), res[1][k].float()
            print("   ", k, tuple(a.shape), "max diff %.3g  n_diff %d" % ((a - b).abs().max().item(), int((a != b).sum())))
