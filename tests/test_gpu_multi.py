"""Multi-GPU tests (-m gpu, need >= 2 CUDA devices; skipped on a single-GPU box): the product kernels under NCCL / NVLink.

One process per GPU (``torch.multiprocessing.spawn``, NCCL backend, rendezvous on 127.0.0.1), rows sharded by rank with
``shard_bounds``:
  * Mahalanobis: ``accumulate(bf16 tensor-core kernel) -> finalize (NCCL all-reduce of the 20 512-double statistics) -> score``
    and ``auroc_fpr95`` (all-reduced key range + histograms) against the single-process float64 spec;
  * sharded similarity / sigmoid loss of the full batch: the peer-memory kernel (B tiles read out of the other ranks' HBM over
    NVLink inside the GEMM) and the NCCL all-gather baseline against the oracle on the gathered batch.
Run here with ``gpurun --gpus 2``; the outcome is recorded in profiles/multi_gpu_tests_r2.txt.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import crossmodal_imu_video_ood_har_b200 as cm
    from crossmodal_imu_video_ood_har_b200.sharded import ShardedSimilarity
    from crossmodal_imu_video_ood_har_b200.models import operand_image
    from oracle import weights as W
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def stage(msg):                      # progress trail on stderr: a hang is attributable from the log of a timed-out run
        import sys
        print(f"[multi rank {rank}] {msg}", file=sys.stderr, flush=True)
    try:
        res = {}
        stage("process group up")
        # ---- Mahalanobis fit across ranks with the tensor-core accumulate kernel
        feats, labels = W.class_features(7, 40003)
        lo, hi = cm.shard_bounds(len(feats), rank, world)
        m = cm.MahalanobisOOD(32, dev, ridge=1e-3)
        m.accumulate(torch.from_numpy(feats[lo:hi]).to(dev), torch.from_numpy(labels[lo:hi]).to(dev), precision="bf16")
        stage("accumulated")
        m.finalize()                                                       # NCCL all-reduce inside
        stage("finalized")
        res["mean"], res["whiten"], res["count"] = m.fit_["mean"], m.fit_["whiten"], m.fit_["count"]
        q, ql = W.class_features(8, 30000, ood_fraction=0.4)
        lo2, hi2 = cm.shard_bounds(len(q), rank, world)
        sc = m.score(torch.from_numpy(q[lo2:hi2]).to(dev), precision="bf16")
        is_ood = torch.from_numpy(ql[lo2:hi2] < 0).to(dev)
        r = cm.auroc_fpr95(sc[~is_ood].contiguous(), sc[is_ood].contiguous())   # all-reduced key range + histograms
        res["auroc"], res["fpr"], res["bound"] = r["auroc"], r["fpr"], r["auroc_bound"]
        res["scores"] = sc.cpu().numpy()
        stage("auroc done")
        # ---- sharded similarity: 256 rows per rank, unit-norm embeddings from a shared seed
        rows, dim = 256, 256
        g = torch.Generator().manual_seed(99)
        a = torch.nn.functional.normalize(torch.randn(rows * world, dim, generator=g), dim=1)
        b = torch.nn.functional.normalize(torch.randn(rows * world, dim, generator=g), dim=1)
        import test_gpu_fused_tail as tail                                # host restatement of the operand-image layout
        a_img = tail.image_of(a[rank * rows:(rank + 1) * rows].to(dev))
        for transport in ("peer", "nccl"):
            ss = ShardedSimilarity(rows, dim, dev, transport=transport)
            stage(f"{transport}: buffers mapped")
            ss.video_image().copy_(tail.image_of(b[rank * rows:(rank + 1) * rows].to(dev)))
            vals = []
            for _ in range(3):                                             # epochs advance, tickets re-arm
                vals.append(float(ss(a_img)))
            torch.cuda.synchronize(dev)
            res[f"loss_{transport}"] = np.array(vals)
            stage(f"{transport}: 3 eager steps done {vals}")
            # captured in a CUDA graph (peer transport only: barriers + similarity are plain kernel launches)
            if transport == "peer":
                s = torch.cuda.Stream(device=dev)
                s.wait_stream(torch.cuda.current_stream(dev))
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=s):
                    ss(a_img)
                ss.loss.zero_()
                gr.replay(); gr.replay()
                torch.cuda.synchronize(dev)
                res["loss_peer_graph"] = float(ss.loss)
                stage("peer: graph replays done")
            dist.barrier(device_ids=[rank])
            ss.close()
            stage(f"{transport}: closed")
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), **res)
    except BaseException:
        # a rank that fails must not sit in destroy_process_group while the others wait in a collective: report and die,
        # mp.spawn then terminates the remaining ranks and raises in the parent
        import traceback
        traceback.print_exc()
        os._exit(1)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_gpu_mahalanobis_auroc_and_sharded_similarity(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices")
    import torch.multiprocessing as mp
    from oracle import ood_spec, oracle, weights as W
    import crossmodal_imu_video_ood_har_b200 as cm
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    feats, labels = W.class_features(7, 40003)
    fit = ood_spec.mahalanobis_fit(feats, labels, 32, ridge=1e-3)
    for g in got:                                                          # identical state on every rank
        np.testing.assert_array_equal(g["whiten"], got[0]["whiten"])
        np.testing.assert_allclose(g["count"], fit["count"])
        np.testing.assert_allclose(g["mean"], fit["mean"], atol=2e-6)
        np.testing.assert_allclose(g["whiten"] @ g["whiten"].T, fit["precision"], rtol=2e-4, atol=2e-5)
    q, ql = W.class_features(8, 30000, ood_fraction=0.4)
    want = ood_spec.mahalanobis_score(q, fit)
    scores = np.concatenate([g["scores"] for g in got])
    assert float(np.abs(scores - want).max() / np.abs(want).max()) < 1e-3
    want_auc, want_fpr = ood_spec.auroc(want[ql >= 0], want[ql < 0]), ood_spec.fpr_at_tpr_fast(want[ql >= 0], want[ql < 0])
    for g in got:
        assert round(float(g["auroc"]), 3) == round(want_auc, 3) and round(float(g["fpr"]), 3) == round(want_fpr, 3)
        assert float(g["auroc"]) == float(got[0]["auroc"])
    # sharded similarity: both transports, every rank, every repetition -> the oracle's full-batch loss
    rows, dim = 256, 256
    gen = torch.Generator().manual_seed(99)
    a = torch.nn.functional.normalize(torch.randn(rows * world, dim, generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(rows * world, dim, generator=gen), dim=1)
    want_loss = float(oracle.sigmoid_contrastive_loss(a.numpy(), b.numpy(), dtype=torch.float64))
    for g in got:
        for key in ("loss_peer", "loss_nccl"):
            assert np.all(np.abs(g[key] - want_loss) < 2e-2 * want_loss), (key, g[key], want_loss)
            assert np.all(g[key] == g[key][0])                            # repetitions are bit-stable
        assert float(g["loss_peer_graph"]) == float(g["loss_peer"][0])
        assert np.array_equal(g["loss_peer"], got[0]["loss_peer"])        # identical bits on every rank
        assert abs(float(g["loss_peer"][0]) - float(g["loss_nccl"][0])) < 1e-12 * want_loss
