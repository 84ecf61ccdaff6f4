"""Console logging with the reference's line formats (``tools/logger.py:3-39``): one logger per actor name."""
from __future__ import annotations

import logging

logging.basicConfig(level=logging.INFO, format="%(asctime)s [%(name)s]: %(levelname)s - %(message)s")


class Logger:
    def __init__(self, actuator: str = "unknown", rank: int | None = None):
        self.logger = logging.getLogger(actuator if rank is None else f"{actuator}@r{rank}")

    def enabled_for_info(self) -> bool:
        return self.logger.isEnabledFor(logging.INFO)

    def debug(self, msg) -> None:
        self.logger.debug(msg)

    def info(self, msg) -> None:
        self.logger.info(msg)

    def warn(self, msg) -> None:
        self.logger.warning(msg)

    warning = warn

    def error(self, msg) -> None:
        self.logger.error(msg)

    def info_train(self, task_name, device, train_cnt, accuracy, loss, current_epoch=0, total_epoch=0) -> None:
        prefix = f"[{current_epoch:0>3d}/{total_epoch:0>3d}] " if current_epoch and total_epoch else ""
        self.logger.info(f"{prefix}Train '{task_name}' on {device} with {train_cnt:,} images, "
                         f"accuracy: {accuracy:.2%}, loss: {loss:.4f}.")

    def info_validation(self, task_name, query_cnt, gallery_cnt, cmc, mAP) -> None:
        r = lambda k: cmc[k] if len(cmc) > k else cmc[-1]  # noqa: E731  (tiny galleries)
        self.logger.info(
            f"Validation '{task_name}' with {query_cnt:,} query images on {gallery_cnt:,} gallery images:\n"
            f"            |- Rank-1 :  {r(0):.2%}\n            |- Rank-3 :  {r(2):.2%}\n"
            f"            |- Rank-5 :  {r(4):.2%}\n            |- Rank-10 : {r(9):.2%}\n"
            f"            |- mean AP : {mAP:.2%}\n            ")
