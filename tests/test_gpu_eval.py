"""Embedding extraction (ref eval/verification.py:238-305) against the oracle's CPU model: originals + flipped copies,
((x/255)-0.5)/0.5, summed, L2-normalised; the overlapping last batch of the reference is reproduced."""
import numpy as np
import pytest
import torch

from gpu_util import assert_close, need_gpu
from oracle.detfill import fill_state_dict_

pytestmark = pytest.mark.gpu


def _net(fp16):
    from msml_b200.backbones import MSML
    net = MSML("iresnet18", "unet", (1, 1, 1, 1), 97, fp16=fp16, header_type="AMArcFace", header_params=(64.0, 0.5, 0.0, 0.0),
               fm_params=(3, 2, "sigmoid", "mul"))
    fill_state_dict_(net)
    return net.cuda().eval()


def _oracle_embeddings(net, imgs_u8):
    from oracle import model_cpu
    sd = {k: v.detach().cpu().float() for k, v in net.state_dict().items()}
    outs = []
    for data in (imgs_u8, torch.flip(imgs_u8, dims=[3])):
        x = ((data.float() / 255) - 0.5) / 0.5
        with torch.no_grad():
            feat, _seg = model_cpu.msml_forward(sd, x, "iresnet18", training=False, fm_params=(3, 2, "sigmoid", "mul"))
        outs.append(feat.double().numpy())
    tot = outs[0] + outs[1]
    return tot / np.linalg.norm(tot, axis=1, keepdims=True), outs


@pytest.mark.parametrize("explicit_flip", [False, True])
def test_extract_embeddings_fp32_matches_oracle(explicit_flip):
    need_gpu()
    from msml_b200.eval import extract_embeddings
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = _net(False)
        g = torch.Generator().manual_seed(5)
        imgs = torch.randint(0, 256, (10, 3, 112, 112), generator=g, dtype=torch.uint8)
        data_list = [imgs.numpy(), torch.flip(imgs, dims=[3]).numpy()] if explicit_flip else [imgs]
        emb, emb_list = extract_embeddings(data_list, net, batch_size=4)      # 10 = 4 + 4 + overlapping tail of 2
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    want, want_list = _oracle_embeddings(net, imgs)
    for a, b in zip(emb_list, want_list):
        assert_close(a, b, 2e-3, atol_frac=2e-3, what="raw embeddings")
    assert_close(emb, want, 2e-3, atol_frac=2e-3, what="normalised flip-sum embeddings")
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-6)
    assert not net.training


def test_extract_embeddings_bf16_and_masks():
    need_gpu()
    from msml_b200.eval import extract_embeddings, random_block_occlusion
    net = _net(True)
    g = torch.Generator(device="cuda").manual_seed(7)
    imgs = torch.randint(0, 256, (8, 3, 112, 112), generator=g, dtype=torch.uint8, device="cuda")
    occ = random_block_occlusion(imgs, 40, 41, generator=g)
    side = int((0.40 * 112 * 112) ** 0.5)                                          # 40 % of the AREA (ref rand_occ.py:45-50): 70 x 70
    assert occ.dtype == torch.uint8 and (occ == 0).sum() >= 8 * 3 * side * side
    black = (occ == 0).all(dim=1)                                                  # the block is inside the frame and square
    assert all(int(b.any(1).sum()) >= side and int(b.any(0).sum()) >= side for b in black)
    emb, _lists, masks = extract_embeddings([occ], net, batch_size=8, return_masks=True)
    want, _ = _oracle_embeddings(net, occ.cpu())
    cos = (emb * want).sum(1)
    assert (cos > 0.995).all(), cos
    assert masks.shape == (8, 112, 112) and set(np.unique(masks)) <= {0, 1}


def test_extract_embeddings_rejects_cpu_model_and_oversized_batch():
    need_gpu()
    from msml_b200.eval import extract_embeddings
    net = _net(False)
    imgs = torch.zeros((2, 3, 112, 112), dtype=torch.uint8)
    with pytest.raises(ValueError):
        extract_embeddings([imgs], net, batch_size=4)
    with pytest.raises(RuntimeError):
        extract_embeddings([imgs], net.cpu(), batch_size=2)
