"""North-star end-to-end check: run_cgs_models (the public API) on a synthetic matrix of a named shape, all
visible GPUs, wall clock from the call to the returned list of CistopicLDAModel (token-list construction,
sweeps, metrics and DataFrames included).  Default = BASELINE.json target T: 20 000 cells x 300 000 regions,
8 models, 500 iterations (< 120 s on 8 x B200)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch
from pycistopic_b200.lda_models import run_cgs_models, evaluate_models
from pycistopic_b200.synth import SHAPES, SyntheticCistopicObject, planted_topic_tokens

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="T")
ap.add_argument("--topics", default="5,10,15,20,25,30,40,50")
ap.add_argument("--n-iter", type=int, default=500)
args = ap.parse_args()
D, W, dens = SHAPES[args.shape]
t = time.time()
cell, region = planted_topic_tokens(D, W, dens, seed=1234, device=torch.device("cuda:0"))
cell, region = cell.cpu().numpy(), region.cpu().numpy()
m = sp.csr_matrix((np.ones(cell.shape[0], np.int32), (region, cell)), shape=(W, D), dtype=np.int32)
m.sort_indices()
torch.cuda.empty_cache()
obj = SyntheticCistopicObject(m, [f"cell_{i}" for i in range(D)], [f"chr1:{i*500}-{i*500+500}" for i in range(W)])
print(f"matrix {args.shape}: {W} regions x {D} cells, nnz {m.nnz} ({time.time()-t:.1f} s to generate)", flush=True)
topics = [int(k) for k in args.topics.split(",")]
n_gpu = torch.cuda.device_count()
t0 = time.time()
models = run_cgs_models(obj, n_topics=topics, n_iter=args.n_iter, random_state=555, alpha=50, alpha_by_topic=True,
                        eta=0.1, eta_by_topic=False)
wall = time.time() - t0
best = evaluate_models(models, return_model=True, plot=False)
print(json.dumps({"shape": args.shape, "cells": D, "regions": W, "tokens": int(m.nnz), "n_topics": topics,
                  "n_iter": args.n_iter, "gpus": n_gpu, "wall_s": round(wall, 2),
                  "token_updates_per_s": m.nnz * args.n_iter * len(topics) / wall,
                  "model_fit_s": {int(mm.n_topic): round(float(mm.parameters.loc["time", "Parameter"]), 2) for mm in models},
                  "loglikelihood": {int(mm.n_topic): float(mm.metrics.loc["Metric", "loglikelihood"]) for mm in models},
                  "selected_n_topic": int(best.n_topic)}), flush=True)
