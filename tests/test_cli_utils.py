"""CLI compatibility, LR schedule, meters, fake data."""
import math

import torch

from vit_10b_fsdp_example_b200.config import ViTConfig, parse_args
from vit_10b_fsdp_example_b200.data import FakeImageNetDataset, build_datasets
from vit_10b_fsdp_example_b200.utils import SmoothedValue, get_warmup_cosine_scheduler

README_CMD = ("--data_dir /datasets/imagenet-1k --ckpt_dir /tmp/vit_fsdp --image_size 224 --patch_size 14 "
              "--embed_dim 5120 --mlp_ratio 4.0 --num_heads 32 --num_blocks 32 --batch_size 1024 --num_epochs 300 "
              "--lr 1e-3 --weight_decay 0.1 --clip_grad_norm 1.0 --warmup_steps 10000 --log_step_interval 20 "
              "--shard_on_cpu").split()


def test_readme_command_line_parses_unchanged():
    cfg = parse_args(README_CMD)
    assert cfg.embed_dim == 5120 and cfg.num_blocks == 32 and cfg.shard_on_cpu and cfg.batch_size == 1024
    assert cfg.grad_ckpt and cfg.reshard_after_forward and not cfg.flatten_parameters and not cfg.run_without_fsdp


def test_defaults_are_the_vit10b_recipe():
    cfg = parse_args([])
    ref = dict(data_dir="/datasets/imagenet-1k", fake_data=False, num_workers=4, ckpt_dir="/tmp/vit_fsdp",
               resume_epoch=0, ckpt_epoch_interval=10, test_epoch_interval=10, log_step_interval=20, image_size=224,
               patch_size=14, embed_dim=5120, num_heads=32, num_blocks=32, mlp_ratio=4.0, pos_dropout=0.0,
               att_dropout=0.0, mlp_dropout=0.0, num_classes=1000, batch_size=1024, num_epochs=300, lr=1e-3,
               weight_decay=0.1, clip_grad_norm=1.0, warmup_steps=10000, grad_ckpt=True, reshard_after_forward=True,
               flatten_parameters=False, run_without_fsdp=False, shard_on_cpu=False)
    for k, v in ref.items():
        assert getattr(cfg, k) == v, k
    flags = parse_args(["--no_grad_ckpt", "--no_reshard_after_forward", "--flatten_parameters", "--run_without_fsdp",
                        "--fake_data"])
    assert not flags.grad_ckpt and not flags.reshard_after_forward and flags.flatten_parameters
    assert flags.run_without_fsdp and flags.fake_data
    v = ViTConfig.from_args(cfg)
    assert v.num_patches == 256 and v.head_dim == 160 and v.hidden_dim == 20480


def test_warmup_cosine_schedule():
    class Opt:
        def __init__(self):
            self.param_groups = [{"lr": 1e-3}]

    opt = Opt()
    sched = get_warmup_cosine_scheduler(opt, warmup_iteration=10, max_iteration=110)
    assert opt.param_groups[0]["lr"] == 0.0
    for _ in range(5):
        sched.step()
    assert abs(opt.param_groups[0]["lr"] - 0.5e-3) < 1e-12
    for _ in range(5):
        sched.step()
    assert abs(opt.param_groups[0]["lr"] - 1e-3) < 1e-12
    for _ in range(50):
        sched.step()
    assert abs(opt.param_groups[0]["lr"] - 1e-3 * 0.5 * (1 + math.cos(math.pi * 0.5))) < 1e-12
    state = sched.state_dict()
    opt2 = Opt()
    s2 = get_warmup_cosine_scheduler(opt2, 10, 110)
    s2.load_state_dict(state)
    assert opt2.param_groups[0]["lr"] == opt.param_groups[0]["lr"]


def test_smoothed_value():
    m = SmoothedValue(window_size=3)
    for v in [1.0, 2.0, 3.0, 4.0]:
        m.update(v, batch_size=1)
    assert m.avg == 3.0 and m.median == 3.0 and m.global_avg == 2.5 and m.get_latest() == 4.0


def test_fake_data_pipeline():
    ds = FakeImageNetDataset(32, 100)
    img, label = ds[7]
    assert img.shape == (3, 32, 32) and float(img.abs().sum()) == 0.0 and label == 0 and len(ds) == 100
    cfg = parse_args(["--fake_data", "--image_size", "32", "--batch_size", "8"])
    train_ds, train_loader, train_sampler, val_ds, val_loader, _ = build_datasets(cfg, torch.device("cpu"), 2, 1,
                                                                                 log=lambda *a: None)
    assert len(train_ds) == 1281167 and len(val_ds) == 50000
    assert len(train_loader) == (1281167 // 2) // 4
    data, target = next(iter(train_loader))
    assert data.shape == (4, 3, 32, 32) and target.shape == (4,) and int(target.sum()) == 0


def test_image_folder_pipeline(tmp_path):
    """Real-data path (reference run_vit_training.py:39-56,62-88): ImageFolder(train/val) + the reference's transforms,
    DistributedSampler (drop_last) and a DataLoader, on a generated 2-class image tree."""
    from PIL import Image

    rng = torch.Generator().manual_seed(0)
    for split, per_class in (("train", 6), ("val", 4)):
        for cls in ("n01", "n02"):
            d = tmp_path / split / cls
            d.mkdir(parents=True)
            for i in range(per_class):
                arr = (torch.rand(40, 52, 3, generator=rng) * 255).to(torch.uint8).numpy()
                Image.fromarray(arr).save(d / f"img_{i}.jpeg")
    cfg = parse_args(["--data_dir", str(tmp_path), "--image_size", "32", "--batch_size", "4", "--num_workers", "0"])
    logs = []
    train_ds, train_loader, train_sampler, val_ds, val_loader, val_sampler = build_datasets(
        cfg, torch.device("cpu"), 2, 0, log=logs.append)
    assert "loading images from directory" in logs[0]
    assert len(train_ds) == 12 and len(val_ds) == 8
    assert len(train_sampler) == 6 and len(val_sampler) == 4  # this rank's half
    train_sampler.set_epoch(1)
    batches = list(train_loader)
    assert len(batches) == 3  # 6 samples / local batch 2, drop_last
    data, target = batches[0]
    assert data.shape == (2, 3, 32, 32) and data.dtype == torch.float32 and target.dtype == torch.long
    assert set(int(t) for b in batches for t in b[1]) <= {0, 1}
    vdata, _ = next(iter(val_loader))
    assert vdata.shape == (2, 3, 32, 32)
    # Normalize() was applied: values are not confined to [0, 1]
    assert float(data.min()) < 0.0


def test_backend_resolution(monkeypatch):
    from vit_10b_fsdp_example_b200.train import resolve_backend

    cuda, cpu = torch.device("cuda"), torch.device("cpu")
    auto = parse_args(["--fake_data"])
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    assert resolve_backend(auto, cuda) == "sm100" and resolve_backend(auto, cpu) == "torchdist"
    # a job that spans hosts cannot use the symmetric-memory kernels: fall back to NCCL, refuse an explicit sm100
    monkeypatch.setenv("WORLD_SIZE", "16")
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert resolve_backend(auto, cuda) == "torchdist"
    import pytest
    with pytest.raises(ValueError):
        resolve_backend(parse_args(["--backend", "sm100"]), cuda)
    assert resolve_backend(parse_args(["--backend", "nccl"]), cuda) == "torchdist"


def test_pod_launcher_builds_per_host_commands(capsys):
    """Multi-host fan-out (role of xla_dist in the reference README): node ranks, rendezvous endpoint, env forwarding."""
    from vit_10b_fsdp_example_b200 import pod_launch

    rc = pod_launch.main(["--hosts", "gpu-a,gpu-b", "--nproc-per-node", "8", "--env", "NCCL_DEBUG=WARN", "--env",
                          "FOO=a b", "--workdir", "/srv/vit", "--restart", "--dry-run", "--", "run_vit_training.py",
                          "--fake_data", "--batch_size", "1024"])
    assert rc == 0
    lines = capsys.readouterr().out.strip().splitlines()
    assert len(lines) == 4 and lines[0].startswith("[gpu-a] if [ -f ")
    a, b = lines[1], lines[3]
    for ln, rank in ((a, 0), (b, 1)):
        assert f"--node-rank={rank}" in ln and "--nnodes=2" in ln and "--nproc-per-node=8" in ln
        assert "--master-addr=gpu-a" in ln and "--master-port=29500" in ln
        assert "NCCL_DEBUG=WARN" in ln and "FOO='a b'" in ln and "cd /srv/vit" in ln
        assert ln.endswith("run_vit_training.py --fake_data --batch_size 1024")
    assert "pkill" not in a and "killall" not in a  # restarts stop the recorded PID, never a pattern


def test_runtime_host_helpers_single_process():
    """xm.* replacements (launch.Runtime): master_print, mesh_reduce, step closures, memory info."""
    from vit_10b_fsdp_example_b200.launch import Runtime

    rt = Runtime(rank=0, world=1, local_rank=0, device=torch.device("cpu"))
    assert rt.mesh_reduce("tag", 3, sum) == 3 and rt.mesh_reduce("tag", 2.5, max) == 2.5
    seen = []
    rt.add_step_closure(lambda a, b: seen.append((a, b)), args=(1, 2))
    rt.add_step_closure(lambda a: seen.append(a), args=("x",))
    assert seen == []           # closures run after the step, not when they are added
    rt.run_step_closures()
    assert seen == [(1, 2), "x"]
    rt.run_step_closures()
    assert seen == [(1, 2), "x"]  # each closure runs once
    info = rt.get_memory_info()
    assert info["kb_total"] >= info["kb_free"] >= 0
    rt.rendezvous("no-op on one process")
