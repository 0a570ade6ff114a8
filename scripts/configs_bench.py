"""GPU box: device-timed throughput of the OTHER BASELINE.json configurations on one B200 (the headline metric,
C3, is bench.py's).  Same method as bench.py: the whole trajectory of one per-GPU batch as one CUDA graph, CUDA
events, 256 MiB L2 flush between iterations.   python scripts/configs_bench.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rangeldm_b200 as R
import bench
from rangeldm_b200.pipelines import FusedSampler, make_pos_encoding

UNET_C2 = dict(sample_size=[1024, 64], in_channels=3, out_channels=2, layers_per_block=2,
               block_out_channels=[128, 128, 256, 256, 512, 512],
               down_block_types=["DownBlock2D"] * 4 + ["AttnDownBlock2D", "DownBlock2D"],
               up_block_types=["UpBlock2D", "AttnUpBlock2D"] + ["UpBlock2D"] * 4)


dev = torch.device("cuda:0")


def unet(cfg):
    torch.manual_seed(0)
    u = R.UNet2DModel(**cfg)
    R.replace_down(u); R.replace_conv(u)
    return u.to(dev)


def vae():
    torch.manual_seed(1)
    v = R.AutoencoderKL(**bench.VAE_KITTI)
    v.quant_conv = torch.nn.Identity(); v.post_quant_conv = torch.nn.Identity()
    R.replace_down(v); R.replace_conv(v); R.replace_attn(v)
    return v.to(dev)


def timed(s, n=3):
    for _ in range(2):
        s.replay()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def report(name, batch, ms, gflop_per_image):
    print(json.dumps({"config": name, "per_gpu_batch": batch, "ms_per_batch": round(ms, 2),
                      "images_per_s": round(batch / ms * 1e3, 2),
                      "achieved_tflops": round(batch * gflop_per_image / ms, 1)}), flush=True)


flush = None


def main():
    global flush
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # C2: RangeDM pixel space, 50-step DDIM, batch 1, no VAE (BASELINE configs[1]); 498.2 GFLOP per UNet forward
    u2 = unet(UNET_C2)
    sch = R.DDIMScheduler(clip_sample=False); sch.set_timesteps(50)
    s = FusedSampler(u2, sch, None, 1, 1)
    s.load(torch.randn(1, 2, 1024, 64, device=dev), make_pos_encoding(1, 1024, 64, dev))
    report("C2 RangeDM pixel 64x1024, 50-step DDIM", 1, timed(s), 50 * 498.2)
    del s, u2
    torch.cuda.empty_cache()
    # C4: nuScenes latent 256x8 -> 32x1024 image, 20-step DPM-Solver, per-GPU batch 16 (BASELINE configs[3])
    u4 = unet(dict(bench.UNET_C3, sample_size=[256, 8])); v = vae()
    sch = R.DPMSolverMultistepScheduler(timestep_spacing="leading"); sch.set_timesteps(20)
    s = FusedSampler(u4, sch, v, 16, 1)
    s.load(torch.randn(16, 4, 256, 8, device=dev), make_pos_encoding(16, 256, 8, dev))
    report("C4 RangeLDM nuScenes 32x1024, 20-step DPM-Solver", 16, timed(s), 20 * 16.28 + 78.73)
    del s, u4
    torch.cuda.empty_cache()
    # C5: conditional upsampling, 4 + 8 channels in, per-GPU batch 4 (BASELINE configs[4])
    u5 = unet(dict(bench.UNET_C3, in_channels=12))
    s = FusedSampler(u5, sch, v, 4, 8)
    s.load(torch.randn(4, 4, 256, 16, device=dev), torch.randn(4, 8, 256, 16, device=dev))
    report("C5 conditional upsample KITTI-360, 20-step DPM-Solver", 4, timed(s), 20 * 34.14 + 157.46)



if __name__ == "__main__":
    main()
