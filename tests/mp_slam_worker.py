"""worker of tests/test_parallel_gpu.py: one rank of a world-size-N SLAM run (torchrun, NCCL), short half-resolution sequence;
rank 0 writes the summary (loss, Gaussian count over all ranks, PSNR of the evaluation renders) as JSON to argv[1]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n_frames, scale, world, rank, local, track=0, split=0):
    import hashlib
    import numpy as np
    import torch
    from gps_slam_b200 import engine as E, slam, synthetic as syn
    E.load_library()
    dev = torch.device("cuda", local)
    intr = syn.intrinsics("replica", scale)
    poses = syn.trajectory(n_frames)
    H, W = intr["height"], intr["width"]
    rgba = torch.empty((n_frames, H, W, 4), dtype=torch.uint8, device=dev)
    depth = torch.empty((n_frames, H, W), dtype=torch.int16, device=dev)
    for i in range(n_frames):
        rgba[i], depth[i] = syn.render_frame(poses[i], intr, device=dev)
    stream = torch.cuda.Stream(device=dev)
    if split:
        # functional split (gps_slam_b200/split.py): rank 0 = the TSDF side, ranks 1.. = Gaussian shards; results live on rank 1
        from gps_slam_b200 import split as split_mod
        pipe = split_mod.SplitSlamPipeline(intr, device=local, stream=stream, rank=rank, world=world, gs_capacity=1 << 19)
    else:
        pipe = slam.SlamPipeline(intr, mode="train", device=local, stream=stream, rank=rank, world=world, gs_capacity=1 << 19,
                                 use_gt_pose=track == 0, tracker=track or 1)
    with torch.cuda.stream(stream):
        for f in range(n_frames):
            pipe.process_frame(f, rgba, depth, poses, True)
        pipe.end_of_step(False)
        pipe.flush_readback()
        rgb, dep, alpha = torch.empty((H, W, 3), device=dev), torch.empty((H, W), device=dev), torch.empty((H, W), device=dev)
        ps = []
        for i in range(0, n_frames, 10):
            pipe.render_eval(poses[i], rgb, dep, alpha)
            torch.cuda.synchronize()
            mse = float(((rgb.clamp(0, 1) - rgba[i][..., :3].float() / 255.0) ** 2).mean())
            ps.append(20.0 * np.log10(1.0 / np.sqrt(mse)))
    st = pipe.stats()
    if split:
        import torch.distributed as dist
        mine = dict(world=world, loss=pipe.last_loss, gaussians=st["gaussians"], psnr=ps, overflow=st["overflow_flags"],
                    gaussians_this_rank=st["gaussians_this_rank"], mailbox_errors=st["mailbox_errors"], cycles=st["cycles"])
        got = [None] * world
        dist.all_gather_object(got, mine)
        pipe.close()
        return dict(got[1], ranks=[g["gaussians_this_rank"] for g in got])
    # TSDF side: replicated state and the last free-view render (every rank holds all rows; sharded: marched by all ranks, voxels read
    # from their owners over NVLink) -- must be bit-identical to the single-GPU run when the poses are given
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()   # noqa: E731
    tsdf = dict(hash_table=sha(pipe.tsdf.hash_entries()), visible_ids=sha(pipe.tsdf.visible_ids()), free_vertex=sha(pipe.tsdf.raycast(live=False)),
                free_image=sha(pipe.tsdf.free_image()), shard_error=pipe.tsdf.shard_error(), sharded=bool(pipe.tsdf_sharded),
                owned_visible=pipe.tsdf.counter(6) if pipe.tsdf_sharded else pipe.tsdf.counter(2), visible=pipe.tsdf.counter(2))
    out = dict(world=world, loss=pipe.last_loss, gaussians=st["gaussians"], psnr=ps, overflow=st["overflow_flags"],
               gaussians_this_rank=st.get("gaussians_this_rank", st["gaussians"]), tsdf=tsdf, tracking=pipe.tracking_stats(poses, n_frames))
    pipe.close()
    return out


if __name__ == "__main__":
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run(int(sys.argv[2]), float(sys.argv[3]), world, rank, local, int(sys.argv[4]) if len(sys.argv) > 4 else 0,
              int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    if rank == 0:
        with open(sys.argv[1], "w") as f:
            json.dump(res, f)
    dist.destroy_process_group()
