"""Grad-CAM visualisation (``analyse/visualize.py``) without the ``pytorch_grad_cam`` / ``cv2`` dependencies:
a ~20-line Grad-CAM over ``model.net.base.layer4[-1]`` (or any module) that returns the heat-map tensors and can
write blended JPEGs with PIL."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def grad_cam(model: torch.nn.Module, layer: torch.nn.Module, images: torch.Tensor,
             target: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Grad-CAM heat-maps ``[B, H, W]`` in [0, 1]. The score is the norm of the embedding when ``target`` is None
    (ReID models return features in eval mode) or the selected class logit otherwise."""
    acts, grads = [], []
    h1 = layer.register_forward_hook(lambda m, i, o: acts.append(o))
    h2 = layer.register_full_backward_hook(lambda m, gi, go: grads.append(go[0]))
    try:
        model.zero_grad(set_to_none=True)
        out = model(images)
        if isinstance(out, tuple):
            out = out[0]
        score = out.norm(dim=1).sum() if target is None else out.gather(1, target.view(-1, 1)).sum()
        score.backward()
        a, g = acts[-1], grads[-1]
        w = g.mean(dim=(2, 3), keepdim=True)
        cam = F.relu((w * a).sum(1, keepdim=True))
        cam = F.interpolate(cam, size=images.shape[-2:], mode="bilinear", align_corners=False).squeeze(1)
        cam = cam - cam.amin(dim=(1, 2), keepdim=True)
        return (cam / cam.amax(dim=(1, 2), keepdim=True).clamp_min(1e-12)).detach()
    finally:
        h1.remove(); h2.remove()


def visualize_models(models: Dict[str, torch.nn.Module], samples: Dict[int, Sequence[str]], save_dir: str,
                     size=(256, 128), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> List[str]:
    """For each ``{method_name: ModelModule}`` and ``{class_id: [image paths]}`` write ``{class}_{i}_{method}.jpg``."""
    import numpy as np
    from PIL import Image
    os.makedirs(save_dir, exist_ok=True)
    written = []
    m_t, s_t = torch.tensor(mean).view(1, 3, 1, 1), torch.tensor(std).view(1, 3, 1, 1)
    for name, model in models.items():
        model.eval()
        for p in model.parameters():
            p.requires_grad_(True)
        layer = model.net.base.layer4[-1] if hasattr(model.net.base, "layer4") else model.net.base.layers[-1]
        dev = next(model.parameters()).device
        for cls, paths in samples.items():
            for i, path in enumerate(paths):
                img = Image.open(path).convert("RGB").resize((size[1], size[0]))
                rgb = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
                x = ((rgb - m_t) / s_t).to(dev).requires_grad_(True)
                cam = grad_cam(model.net, layer, x).cpu()[0]
                heat = torch.stack([cam, torch.zeros_like(cam), 1 - cam], 0)             # red = important
                blend = (0.5 * rgb[0] + 0.5 * heat).clamp(0, 1)
                out = os.path.join(save_dir, f"{cls}_{i}_{name}.jpg")
                Image.fromarray((blend.permute(1, 2, 0).numpy() * 255).astype("uint8")).save(out)
                written.append(out)
    return written
