"""Time kernel variants of the fused path on one GPU (development tool; not part of the product or the bench).

    python tools/tune_embed.py [config2|config3|config1] [--variants 4:4,8:3,...]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import scone_b200 as sb  # noqa: E402
from scone_b200.utils import synthetic as S  # noqa: E402


def graph_time(fn, steps=20, reps=5):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for k in range(3):
            fn(k)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for k in range(steps):
                fn(k)
        g.replay()
        stream.synchronize()
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.replay()
            e1.record(stream)
            stream.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
    return best


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "config2"
    variants = ["-1", "1:0:3:6:3:64", "1:0:4:8:2:100", "1:0:4:8:2:64", "1:0:4:4:3:70", "1:0:6:6:2:100", "1:0:8:8:1:200", "1:0:12:12:1:200",
                "1:0:6:10:2:100", "1:0:4:12:2:100", "1:0:8:16:1:200", "1:0:12:20:1:200", "1:0:6:12:1:200", "1:0:3:5:3:70", "1:0:2:6:3:70"]
    for a in sys.argv[1:]:
        if a.startswith("--variants="):
            variants = a.split("=", 1)[1].split(",")
    if name.startswith("custom:"):      # custom:D:quant:max_n:N:B:L:V
        _, D_, q_, mn_, N_, B_, L_, V_ = name.split(":")
        w = dict(N=int(N_), D=int(D_), V=int(V_), max_n=int(mn_), quant=q_, B=int(B_), L=int(L_), desc=name)
    else:
        w = bench.WORKLOADS[name]
    dev = torch.device("cuda", 0)
    B, L, D, N, V = w["B"], w["L"], w["D"], w["N"], w["V"]
    T = B * L
    toks, lens, longest = S.make_vocab_device(N, w["max_n"], V, seed=0, device=dev, return_longest=True)
    lf = float(os.environ.get("LF", "0"))
    index = sb.FGramIndex(toks, lens, load_factor=lf)
    print(json.dumps({"load_factor": lf, "index_MB": index.bytes / 1e6, "max_probe": index.max_probe, "slot_bytes": index.slot_bytes}))
    table = sb.CacheTable(N, D, w["quant"], device=dev, align=int(os.environ.get("ALIGN", "32")))
    print(json.dumps({"row_stride": table.row_stride}))
    S.fill_table_device(table, seed=2)
    base = S.make_base_device(V, D, torch.bfloat16, seed=3, device=dev)
    batches = [S.make_stream_device(toks, lens, B, L, V, seed=100 + k, p_plant=1.0, pick_ids=longest) for k in range(8)]
    out = torch.empty((B, L, D), dtype=torch.bfloat16, device=dev)
    out_id = torch.empty((B, L), dtype=torch.int32, device=dev)
    out_len = torch.empty((B, L), dtype=torch.uint8, device=dev)
    sb.embed_forward(index, table, base, batches[0], out=out, out_id=out_id, out_len=out_len)
    hit = float((out_id >= 0).float().mean().item())
    probes = bin(index.len_mask).count("1")
    bpt = bench.bytes_per_token(w, hit, probes, slot_bytes=index.slot_bytes)
    peak, _ = bench.measured_peak_hbm()
    res = {"workload": name, "hit": hit, "bytes_per_token": bpt, "peak": peak, "rows": []}

    def report(label, ms):
        gbs = bpt * T / (ms * 1e-3) / 1e9
        row = {"variant": label, "us": ms * 1e3, "Mtok_s": T / ms / 1e3, "GBs": gbs, "frac": gbs / peak}
        res["rows"].append(row)
        print(json.dumps(row), flush=True)

    # reference points: a plain copy of the same number of bytes, and the two halves separately
    nbytes = int(bpt * T)
    a = torch.empty(nbytes // 2, dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    report("torch_copy_same_bytes(uint8)", graph_time(lambda k: b.copy_(a)))
    a4, b4 = a[: (nbytes // 2) // 16 * 16].view(torch.float32), b[: (nbytes // 2) // 16 * 16].view(torch.float32)
    report("torch_copy_same_bytes(float32, vectorised)", graph_time(lambda k: b4.copy_(a4)))
    big_a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
    big_b = torch.empty_like(big_a)
    ms = graph_time(lambda k: big_b.copy_(big_a), steps=3, reps=3)
    print(json.dumps({"variant": "torch_copy_4GiB_traffic(bf16)", "us": ms * 1e3, "GBs": 2 * big_a.numel() * 2 / (ms * 1e-3) / 1e9}), flush=True)
    del big_a, big_b
    report("lookup_only", graph_time(lambda k: index.lookup(batches[k % 8])))
    fid, _ = index.lookup(batches[0])
    fids = [index.lookup(batches[k])[0] for k in range(8)]
    report("gather_only(ids resolved)", graph_time(lambda k: sb.embed_gather(table, base, batches[k % 8], fids[k % 8], out=out)))
    for v in variants:
        os.environ["SCONE_EMBED_VARIANT"] = v
        report("fused kind:U:NM:NG:MINB:KB=" + v, graph_time(lambda k: sb.embed_forward(index, table, base, batches[k % 8], out=out, out_id=out_id, out_len=out_len)))
    os.environ.pop("SCONE_EMBED_VARIANT", None)
    pos = S.make_base_device(L, D, torch.bfloat16, seed=5, device=dev)
    report("fused + wpe[position] add", graph_time(lambda k: sb.embed_forward(index, table, base, batches[k % 8], pos_emb=pos, out=out, out_id=out_id, out_len=out_len)))
    report("fused, combine=add (wte row + f-gram row)", graph_time(lambda k: sb.embed_forward(index, table, base, batches[k % 8], out=out, out_id=out_id, out_len=out_len, combine="add")))
    report("fused, combine=add + wpe[position] add", graph_time(lambda k: sb.embed_forward(index, table, base, batches[k % 8], pos_emb=pos, out=out, out_id=out_id, out_len=out_len, combine="add")))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/tune_" + name.replace(":", "_") + ".json", "w"), indent=1)


if __name__ == "__main__":
    main()
