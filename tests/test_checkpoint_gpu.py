"""SURVEY 8f row 4 on the device: asynchronous checkpoints (pinned staging + side stream + writer thread) of CUDA tensors."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_async_writer_cuda_sources_may_be_freed_and_overwritten(tmp_path):
    """The caller drops its clones right after save(): the side-stream copies must still read the ORIGINAL values (record_stream),
    even when the allocator hands the same blocks to new tensors that are overwritten on the compute stream."""
    from ct_clip_b200.checkpoint import AsyncCheckpointWriter
    w = AsyncCheckpointWriter()
    ref = {f"t{i}": torch.full((1 << 22,), float(i), device="cuda") for i in range(8)}        # 8 x 16 MB
    src = {k: v.clone() for k, v in ref.items()}
    w.save(src, tmp_path / "a.pt")
    del src                                            # clones go back to the caching allocator immediately
    junk = [torch.full((1 << 22,), -7.0, device="cuda") for _ in range(16)]      # likely to land on the freed blocks
    for j in junk:
        j.mul_(2.0)
    w.wait()
    got = torch.load(tmp_path / "a.pt")
    for k, v in ref.items():
        assert torch.equal(got[k], v.cpu()), k


def test_trainer_async_checkpoint_matches_parameters_of_that_step(tmp_path):
    from transformers import BertConfig, BertModel

    from ct_clip_b200 import CTCLIP, CTViT
    from ct_clip_b200.data import SyntheticCTReportDataset
    from ct_clip_b200.trainer import CTClipTrainer
    torch.manual_seed(0)
    vit = CTViT(dim=512, codebook_size=256, image_size=32, patch_size=16, temporal_patch_size=4, spatial_depth=1, temporal_depth=1,
                dim_head=32, heads=8)
    bert = BertModel(BertConfig(num_hidden_layers=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    clip = CTCLIP(image_encoder=vit, text_encoder=bert, dim_text=768, dim_image=4 * 512, dim_latent=128)
    ds = SyntheticCTReportDataset(64, frames=8, image=32, n_text=16)
    tr = CTClipTrainer(clip, num_train_steps=3, batch_size=2, train_dataset=ds, num_workers=0, save_model_every=1, lr=1e-3,
                       results_folder=str(tmp_path), async_checkpoints=True)
    tr.print = lambda m: None
    snaps = []
    for _ in range(3):
        snaps.append({k: v.detach().clone().cpu() for k, v in tr.CTClip.state_dict().items()})   # state BEFORE the step = what step i saves
        tr.train_step()
    tr._writer.wait()
    # the checkpoint written in train_step i holds the state AFTER the optimiser update of step i
    after = {k: v.detach().cpu() for k, v in tr.CTClip.state_dict().items()}
    last = torch.load(tmp_path / "CTClip.2.pt")
    assert set(last.keys()) == set(after.keys())
    for k in after:
        assert torch.equal(last[k], after[k]), k
    first = torch.load(tmp_path / "CTClip.0.pt")
    changed = sum(int(not torch.equal(first[k], snaps[0][k])) for k in first if first[k].is_floating_point())
    assert changed > 10           # step 0's file is the post-update state, not the initial weights
