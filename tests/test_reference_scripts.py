"""Drop-in boundary: the reference's launch scripts against THIS package.

CPU part (this file's unmarked tests; needs the reference checkout, present in the build container only): the
UNMODIFIED ``scripts/eval.py`` and ``scripts/finetune.py`` are executed as modules with the product ``cs_vit`` on
the path - every ``from cs_vit... import ...`` of ref:scripts/eval.py:19-22 / ref:scripts/finetune.py:19-23 must
resolve - and eval.py's own ``setup()`` builds its data loader and its ``Poser`` from a reference-format config.
Test infrastructure only: ``h5py`` (not installed here) is stubbed for the import, ``Module.to`` /
``DistributedDataParallel`` are patched because this container has no GPU.

GPU part (``-m gpu``, no reference needed): the evaluation flow of ref:scripts/eval.py:204-317 restated on the
product's pieces - data set shim, DistributedSampler, DDP wrap, ``model.module.predict_batch``, reprojection,
rank-0 gather, result arrays - checked against a direct ``predict_batch``.
"""
import importlib.util
import json
import os
import sys
import types

import pytest
import torch

from helpers import backbone_dir

REF = os.environ.get("CSVIT_REFERENCE_ROOT", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")), reason="reference checkout not present")


def _load_script(name):
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules["h5py"] = types.ModuleType("h5py")      # import-time stub: nothing is written in these tests
    spec = importlib.util.spec_from_file_location(f"ref_script_{name}", os.path.join(REF, "scripts", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    keep = os.environ.get("CUDA_LAUNCH_BLOCKING")
    try:
        spec.loader.exec_module(mod)
    finally:      # finetune.py sets CUDA_LAUNCH_BLOCKING=1 at import; do not leak it into the other tests
        if keep is None:
            os.environ.pop("CUDA_LAUNCH_BLOCKING", None)
        else:
            os.environ["CUDA_LAUNCH_BLOCKING"] = keep
    return mod


@needs_ref
@pytest.mark.parametrize("script", ["eval", "finetune", "benchmark"])
def test_reference_scripts_import_against_this_package(script):
    import cs_vit
    assert "cs-vit_b200" in cs_vit.__file__, "the product package must be the cs_vit on the path"
    mod = _load_script(script)
    if script != "benchmark":
        for name in ("Poser", "warmup_scheduler", "InterHand26MSeq", "HO3D", "DexYCB", "FinetuneConfig", "move_to_device",
                     "flatten_dict", "wrap_prefix_print", "print_grouped_losses"):
            assert hasattr(mod, name), name
        assert mod.Poser.__module__.startswith("cs_vit.net")


@needs_ref
def test_reference_eval_setup_runs_on_the_shim(tmp_path, monkeypatch):
    """ref:scripts/eval.py:87-201 ``setup()``: data set + DataLoader + Poser(**config) + checkpoint load, unmodified."""
    from cs_vit.net import Poser
    from cs_vit.utils.mano_standin import SyntheticMANO
    ev = _load_script("eval")
    with open(os.path.join(REF, "checkpoints", "debug_ft", "config.json")) as f:
        cfg = ev.FinetuneConfig(**json.load(f))          # a reference-format config file, loaded the script's way
    bdir = backbone_dir("swin_t")
    cfg.update({"backbone": bdir, "img_size": 224, "data": "dexycb", "dexycb_root": "synthetic:12", "batch_size": 4, "seq_len": 1})
    # the checkpoint the script loads: {"merged": state_dict} (ref:scripts/eval.py:153)
    torch.manual_seed(0)
    donor = Poser(bdir, image_size=224, mano_layer=SyntheticMANO(), num_pose_query=cfg.num_joints,
                  spatial_layer_type=cfg.spatial_layer_type, persp_decorate=cfg.persp_decorate)
    ckpt = tmp_path / "ckpt.pt"
    torch.save({"merged": donor.state_dict()}, ckpt)
    cfg.update({"eval_ckpt": str(ckpt)})
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setenv("CSVIT_MANO", "synthetic")
    monkeypatch.setattr(torch.nn.Module, "to", lambda self, *a, **k: self)          # no GPU in this container
    monkeypatch.setattr(ev, "DistributedDataParallel", lambda m, **k: types.SimpleNamespace(module=m, parameters=m.parameters))
    import torch.distributed as dist
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        (_, _, _, dataloader, _, scheduler, _, model) = ev.setup(0, cfg, print)
    finally:
        if created:
            dist.destroy_process_group()
    assert isinstance(model.module, Poser) and not model.module.training
    got = model.module.state_dict()
    common = [k for k in donor.state_dict() if k in got]      # strict=False: the script's temporal options differ from the donor's
    assert len(common) > 400 and all(torch.equal(got[k], donor.state_dict()[k]) for k in common)
    assert len(dataloader) == 3
    batch = ev.InterHand26MSeq.collate_fn([dataloader.dataset[0], dataloader.dataset[1]])
    assert batch["patches"].shape == (2, 1, 3, 224, 224) and len(batch["imgs_path"]) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("phase,T", [("spatial", 1), ("temporal", 3)])
def test_eval_flow_on_the_shim_matches_predict_batch(tmp_path, phase, T):
    """The loop body of ref:scripts/eval.py:258-312 (DDP-wrapped model, per-batch gather to rank 0, last-frame slicing,
    reprojection) on the synthetic DexYCB shim; the gathered ``joint_cam_pred`` must equal ``predict_batch`` run directly."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel
    from torch.utils.data.dataloader import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    from cs_vit.dataset import DexYCB, InterHand26MSeq
    from cs_vit.distributed import gather_eval_results
    from cs_vit.net import Poser
    from cs_vit.synthetic import randomize_head_
    from cs_vit.utils.mano_standin import SyntheticMANO
    from cs_vit.utils.misc import move_to_device

    torch.manual_seed(0)
    model = Poser(backbone_dir("swin_t"), image_size=224, mano_layer=SyntheticMANO(), spatial_layer_type="encoder", persp_decorate="patch",
                  temporal_supervision="realtime" if phase == "temporal" else "full", temporal_init_method="random", precision="fp16")
    randomize_head_(model)
    model.phase(Poser.TrainingPhase(phase))          # as ref:scripts/eval.py:150 does with cfg.phase
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        model.to(0)
        model.eval()
        ddp = DistributedDataParallel(model, device_ids=[0], output_device=0, find_unused_parameters=False)
        dataset = DexYCB(root="synthetic:10", num_frames=T, protocol="s1", data_split="test", img_size=224, expansion_ratio=1.25)
        loader = DataLoader(dataset, batch_size=4, pin_memory=False, drop_last=False, num_workers=0,
                            sampler=DistributedSampler(dataset, shuffle=False, drop_last=False), collate_fn=InterHand26MSeq.collate_fn)
        rows = {"img_paths": [], "joint_cam_gt": [], "joint_cam_pred": [], "joint_reproj_gt": [], "joint_reproj_pred": []}
        for batch in loader:
            batch = move_to_device(batch, torch.device("cuda:0"))
            with torch.inference_mode():
                predict = ddp.module.predict_batch(img_tensor=batch["patches"], square_bboxes=batch["square_bboxes"],
                                                   timestamp=batch["timestamp"], focal=batch["focal"], princpt=batch["princpt"])
            jc = predict["joint_cam"]
            u = batch["focal"][..., :1] * jc[..., 0] + batch["princpt"][..., :1] * jc[..., 2]
            v = batch["focal"][..., 1:] * jc[..., 1] + batch["princpt"][..., 1:] * jc[..., 2]
            reproj = (torch.stack([u, v], dim=-1) / jc[..., -1:])[:, -1]
            tensors, paths = gather_eval_results({"joint_cam_gt": batch["joint_cam"][:, -1], "joint_cam_pred": jc[:, -1],
                                                  "joint_reproj_gt": batch["joint_img"][:, -1], "joint_reproj_pred": reproj},
                                                 [p[-1] for p in batch["imgs_path"]])
            rows["img_paths"] += paths
            for k, val in tensors.items():
                rows[k].append(val.float().cpu())
    finally:
        if created:
            dist.destroy_process_group()
    pred = torch.cat(rows["joint_cam_pred"])
    assert pred.shape == (10, 21, 3) and len(rows["img_paths"]) == 10
    assert rows["img_paths"][3] == f"synthetic_dexycb/00000003/{T - 1:04d}.jpg"
    # direct: the same 10 clips in one batch
    full = InterHand26MSeq.collate_fn([dataset[i] for i in range(10)])
    dev = {k: v.cuda() for k, v in full.items() if torch.is_tensor(v)}
    with torch.inference_mode():
        want = model.predict_batch(dev["patches"], dev["square_bboxes"], dev["timestamp"], dev["focal"], dev["princpt"])["joint_cam"][:, -1]
    assert torch.allclose(pred, want.float().cpu(), rtol=1e-5, atol=1e-3), (pred - want.float().cpu()).abs().max()
    assert torch.isfinite(torch.cat(rows["joint_reproj_pred"])).all()
    assert torch.equal(torch.cat(rows["joint_cam_gt"]), full["joint_cam"][:, -1])


@pytest.mark.gpu
def test_own_allreduce_kernel_two_ranks():
    """csvit_allreduce_f32 (symmetric-memory allreduce of the finetune gradients) against NCCL on 2 GPUs: tools/multi_gpu_check.py
    under torchrun.  Needs two devices; the driver's single-GPU test box skips it (the 8-GPU record is profiles/r2_multi_gpu_check_*.txt)."""
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(root, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
