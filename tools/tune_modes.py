"""Time the fused path's modes (replace / +wpe / combine=add / add+wpe) under environment overrides of a SCONE_TUNE build
(development tool).

    python tools/tune_modes.py config2 "mode;ENV=val,ENV=val" ...

mode in {replace, pos, add, addpos}; STABLE=1 among the overrides passes SCONE_EMBED_INPUTS_STABLE (SCONE_NO_EARLY=1 makes the
library ignore it: read once at load time, so use STABLE for A/Bs inside one process).  Overrides understood by the SCONE_TUNE build: SCONE_EMBED_VARIANT=kind:U:NM:NG:MINB:KB,
SCONE_EMBED_P (lanes per position), SCONE_STAGGER_NS / SCONE_STAGGER_CTA_NS, SCONE_HINT (what-if: perfect pre-filter).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import scone_b200 as sb  # noqa: E402
from scone_b200.utils import synthetic as S  # noqa: E402
from tune_embed import graph_time  # noqa: E402

KNOBS = ("SCONE_HINT", "SCONE_EMBED_VARIANT", "SCONE_EMBED_P", "SCONE_STAGGER_NS", "SCONE_STAGGER_CTA_NS", "SCONE_EMBED_PIPE", "SCONE_BASE_POLICY")


def main():
    if os.environ.get("AB_LIB"):          # A/B against another build of the library (development only)
        from scone_b200 import _lib
        _lib.LIB_PATH = os.path.abspath(os.environ["AB_LIB"])
    name = sys.argv[1]
    if name.startswith("custom:"):      # custom:D:quant:max_n:N:B:L:V (as tools/tune_embed.py)
        _, D_, q_, mn_, N_, B_, L_, V_ = name.split(":")
        w = dict(N=int(N_), D=int(D_), V=int(V_), max_n=int(mn_), quant=q_, B=int(B_), L=int(L_), desc=name)
    else:
        w = bench.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    B, L, D, N, V = w["B"], w["L"], w["D"], w["N"], w["V"]
    T = B * L
    toks, lens, longest = S.make_vocab_device(N, w["max_n"], V, seed=0, device=dev, return_longest=True)
    index = sb.FGramIndex(toks, lens, load_factor=float(os.environ.get("LF", "0")))
    print(json.dumps({"LF": os.environ.get("LF", "default"), "index_MB": index.bytes / 1e6, "max_probe": index.max_probe}))
    table = sb.CacheTable(N, D, w["quant"], device=dev)
    S.fill_table_device(table, seed=2)
    base = S.make_base_device(V, D, torch.bfloat16, seed=3, device=dev)
    pos = S.make_base_device(L, D, torch.bfloat16, seed=5, device=dev)
    all_ids = torch.stack([S.make_stream_device(toks, lens, B, L, V, seed=100 + k, p_plant=1.0, pick_ids=longest) for k in range(8)])
    batches = [all_ids[k] for k in range(8)]
    out = torch.empty((B, L, D), dtype=torch.bfloat16, device=dev)
    out_id = torch.empty((B, L), dtype=torch.int32, device=dev)
    out_len = torch.empty((B, L), dtype=torch.uint8, device=dev)
    ref = {}
    # SCONE_HINT=1 (what-if: a perfect pre-filter): the true match lengths of all eight batches, handed to the TUNE build
    from scone_b200 import _lib
    all_len = torch.stack([index.lookup(b)[1] for b in batches]).contiguous()
    if hasattr(_lib.load(), "scone_debug_set_hint"):
        import ctypes as C
        _lib.load().scone_debug_set_hint(C.c_void_p(all_ids.data_ptr()), C.c_void_p(all_len.data_ptr()), C.c_int64(all_ids.numel()))
    for spec in sys.argv[2:]:
        mode, _, envs = spec.partition(";")
        for k in KNOBS:
            os.environ.pop(k, None)
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            os.environ[k] = v
        stable = os.environ.pop("STABLE", "0") == "1"
        kw = dict(pos_emb=pos if mode in ("pos", "addpos") else None, combine="add" if mode in ("add", "addpos") else "replace",
                  inputs_stable=stable)
        fn = lambda k: sb.embed_forward(index, table, base, batches[k % 8], out=out, out_id=out_id, out_len=out_len, **kw)  # noqa: E731
        ms = graph_time(fn)
        # every override must reproduce the default build's bits for the same mode
        fn(0)
        torch.cuda.synchronize()
        digest = (int(out.view(torch.int16).to(torch.int64).sum().item()), int(out_id.to(torch.int64).sum().item()))
        same = ref.setdefault(mode, digest) == digest
        print(json.dumps({"workload": name, "mode": mode, "env": envs, "us": ms * 1e3, "Mtok_s": T / ms / 1e3, "same_bits_as_first": same}), flush=True)


if __name__ == "__main__":
    main()
