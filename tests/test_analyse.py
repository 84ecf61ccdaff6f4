"""Offline analysis over the experiment-log JSON schema (``analyse/accuracy.py``, ``analyse/forgetting.py``)."""
import json

from flpr_b200.analyse import load_logs
from flpr_b200.analyse.accuracy import accuracy_curves, accuracy_on_round, merged_curve
from flpr_b200.analyse.forgetting import forgetting_curves, forgetting_on_round


def _log():
    data = {}
    for c in range(2):
        rounds = {}
        for r in (0, 10, 20):
            rounds[str(r)] = {f"task-{c}-{t}": {"val_map": 0.2 + 0.02 * r - 0.05 * t * (r == 20),
                                                "val_rank_1": 0.3 + 0.02 * r} for t in range(2)}
        data[f"client-{c}"] = rounds
    return {"config": {"exp_name": "x"}, "data": data}


def test_accuracy_and_forgetting_tables(tmp_path):
    path = tmp_path / "log.json"
    path.write_text(json.dumps(_log()))
    logs = load_logs(str(path))
    acc = accuracy_on_round(logs, 20, "val_map", verbose=False)
    assert acc is not None
    curves = accuracy_curves(logs, "val_map")
    assert set(curves) == {"client-0", "client-1"} and len(curves["client-0"]["task-0-0"]) == 3
    merged = merged_curve(logs, "val_rank_1")
    assert [r for r, _ in merged] == [0, 10, 20] and abs(merged[-1][1] - 0.7) < 1e-9
    forget = forgetting_on_round(logs, 20, "val_map", verbose=False)
    assert forget is not None
    fc = forgetting_curves(logs, "val_map")
    assert fc[0][0] == 0 and fc[-1][0] == 20 and fc[-1][1] >= 0.0


def test_grad_cam_and_blended_images(tmp_path):
    """``analyse/visualize.py``: Grad-CAM over ``net.base.layer4[-1]`` and the blended JPEG writer."""
    import numpy as np
    import torch
    from PIL import Image

    from flpr_b200.analyse.visualize import grad_cam, visualize_models
    from flpr_b200.models import nets
    from flpr_b200.runtime.modules import ModelModule
    torch.manual_seed(0)
    model = ModelModule(nets["resnet18"](num_classes=10, last_stride=1, neck="bnneck")).eval()
    x = torch.randn(2, 3, 64, 32)
    cam = grad_cam(model.net, model.net.base.layer4[-1], x)
    assert cam.shape == (2, 64, 32) and float(cam.min()) >= 0.0 and float(cam.max()) <= 1.0 + 1e-6
    assert float(cam.amax(dim=(1, 2)).min()) > 0.99                    # normalised per image
    paths = []
    for i in range(2):
        p = tmp_path / f"im{i}.jpg"
        Image.fromarray(np.random.RandomState(i).randint(0, 255, (80, 40, 3), dtype=np.uint8)).save(p)
        paths.append(str(p))
    written = visualize_models({"fedstil": model}, {7: paths}, str(tmp_path / "out"), size=(64, 32))
    assert len(written) == 2 and all(Image.open(w).size == (32, 64) for w in written)
