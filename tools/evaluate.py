"""Evaluation CLI of the B200 path: the command line of the reference's ``script/evaluation.py`` (:455-577), batched and sharded.

    python tools/evaluate.py --model_type avsr_cocktail --dataset_name lrs2 --set_id test --synthetic 64
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/evaluate.py --dataset_name AVCocktail --set_id '*' --synthetic 4

Prints what the reference prints (``WER: ...`` / ``WER <chunk_type>: ...`` / ``Average WER ...``) on rank 0.  The reference pulls
its datasets and the checkpoint from the HuggingFace hub; this box has no network, so ``--checkpoint_path`` takes a local
``state_dict`` (torch.save / safetensors with the reference's key names) and ``--manifest`` a local torch-saved list of samples
(``{"video": uint8 [T,1,H,W], "audio": float [n,1], "label": str}`` for lrs2; see ``evaluation.evaluate_avcocktail`` for
AVCocktail).  ``--synthetic N`` generates N random samples (random-init weights decode noise: the WER is ~1, the plumbing is
what is exercised).  Only ``--model_type avsr_cocktail`` is built.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser(description="B200 inference / evaluation for the avsr_cocktail model")
    ap.add_argument("--model_type", default="avsr_cocktail", choices=["avsr_cocktail"])
    ap.add_argument("--dataset_name", default="lrs2", choices=["lrs2", "AVCocktail"])
    ap.add_argument("--set_id", default="test")
    ap.add_argument("--checkpoint_path", default=None, help="local state_dict file (reference key names); default: seeded random init")
    ap.add_argument("--cache_dir", default=None, help="accepted for compatibility (nothing is downloaded)")
    ap.add_argument("--beam_size", type=int, default=3)
    ap.add_argument("--max_length", type=int, default=15, help="seconds per fixed chunk (AVCocktail synthetic chunks)")
    ap.add_argument("--manifest", default=None)
    ap.add_argument("--synthetic", type=int, default=0)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--output_dir_name", default="output")
    args = ap.parse_args()
    from avsr_b200 import evaluation as E
    from avsr_b200 import input_pipeline as P
    from avsr_b200 import synth, text
    from avsr_b200.model import AVSRCocktailB200
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.checkpoint_path:
        if args.checkpoint_path.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(args.checkpoint_path)
        else:
            sd = torch.load(args.checkpoint_path, map_location="cpu")
    else:
        sd = synth.make_state_dict(0)
    token_list = ["<blank>"] + [f"▁u{i}" for i in range(synth.V - 2)] + ["<eos>"]
    units = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "baseline", "_ref", "src", "tokenizer", "spm", "unigram", "unigram5000_units.txt")
    if os.path.exists(units):                                   # the reference's own unit list (spm_tokenizer.py:31-38)
        toks = [l.split()[0] for l in open(units, encoding="utf-8").read().splitlines()]
        token_list = ["<blank>"] + toks + ["<eos>"]
    model = AVSRCocktailB200(sd, device=dev, beam_size=args.beam_size, token_list=token_list)
    ids_to_text, norm = text.make_text_functions(token_list)
    collator = P.DataCollator(device=str(dev))

    def collate(samples):
        b = collator([{"video": s["video"], "audio": s["audio"]} for s in samples])
        return b["videos"], b["audios"], b["video_lengths"].tolist()

    def synth_sample(seed, T):
        g = torch.Generator().manual_seed(seed)
        return {"video": torch.randint(0, 256, (T, 1, 96, 96), generator=g, dtype=torch.uint8),
                "audio": 0.1 * torch.randn(T * 640, 1, generator=g), "label": " ".join(f"W{(seed * 7 + k) % 50}" for k in range(3 + seed % 5))}

    say = print if rank == 0 else (lambda *a, **k: None)
    say(f"Inferring {args.dataset_name}/{args.set_id} sessions using {args.model_type} model")
    if args.dataset_name == "lrs2":
        samples = torch.load(args.manifest) if args.manifest else [synth_sample(i, 25 + (i * 13) % 100) for i in range(max(1, args.synthetic))]
        lengths = [int(s["video"].shape[0]) for s in samples]
        res = E.evaluate_sharded(model, lengths, lambda i: samples[i], references=[s["label"] for s in samples], ids_to_text=ids_to_text,
                                 normalize=norm, device=dev, collate=collate)
        say(f"WER: {res.wer}")
    else:
        if args.manifest:
            videos = torch.load(args.manifest)
        else:
            videos = {}
            chunk_T = args.max_length * 25
            for v in range(max(1, args.synthetic)):
                cues = "".join(f"00:00:{5 * k:02d}.000 --> 00:00:{5 * k + 4:02d}.500\nW{(v + k) % 50} W{(v * 3 + k) % 50}\n\n" for k in range(6))
                entry = {"label": "WEBVTT\n\n" + cues}
                for ci, ct in enumerate(E.CHUNK_TYPES):
                    entry[ct] = []
                    for k in range(2):
                        s = synth_sample(1000 * v + 10 * ci + k, chunk_T)
                        entry[ct].append({"start_time": k * args.max_length, "end_time": (k + 1) * args.max_length, "frames": chunk_T,
                                          "load": (lambda s=s: s)})
                videos[f"video_{v}"] = entry
        if args.set_id != "*":
            videos = {k: v for k, v in videos.items() if k == args.set_id} or videos
        per_video, num_words, average = E.evaluate_avcocktail(model, videos, ids_to_text, normalize=norm, device=dev, collate=collate)
        for set_id in per_video:
            for ct, w in per_video[set_id].items():
                say(f"WER {set_id} {ct}: {w:.4f}")
        for ct, w in average.items():
            say(f"Average WER {ct}: {w:.4f}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
