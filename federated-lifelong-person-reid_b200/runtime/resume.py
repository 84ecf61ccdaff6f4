"""Resume manifest (the reference cannot resume - no round counter, RNG, task-pipeline position,
server registry or token memory is restorable; a rerun starts at round 1 on stale model files).

``save(stage, ...)`` snapshots, per rank, everything a continued run needs beyond the reference-layout checkpoints:

    {ckpt_root}/{exp}/_resume/rank{r}.ckpt
        round, python / torch / per-model CUDA generator states,
        server : registry (``clients`` keys), upload bookkeeping, method extras (FedSTIL token memory, ...)
        clients: counters, task-pipeline position, optimizer lr / scheduler epoch, full model state,
                 method extras (FedSTIL global weight, exemplar generations, task tokens, ...)
        comm   : this rank's slots of every symmetric client buffer (the *last uploads* that the stale-client
                 aggregation of ``methods/fedavg.py:386-397`` keeps using) and the rank buffers

It goes through the :class:`CheckpointStore` like every other snapshot. ``load(stage, ...)`` restores it after
``ExperimentStage.build`` and returns the round to continue from. Method plug-ins extend the picture through
``resume_extra() / load_resume_extra()`` on their ``Model`` / ``Client`` / ``Server`` classes.

**Rank loss / elastic restart** (SURVEY 5.3). A rank that dies takes the job down: its peers' collectives time out
(flag watchdog -> ``NativeError``; gloo / NCCL: a failed collective), ``torchrun --max-restarts N`` restarts the whole
group, and with ``engine_opts.resume`` every rank continues from the manifest. For that the manifests of the ranks must
describe the SAME round even when the crash hits in the middle of a save:

* two generations per rank, ``rank{r}-g0`` / ``rank{r}-g1``; a save overwrites the older one;
* a generation counts only with its commit marker ``rank{r}-g{k}.ckpt.ok`` (``{"round": n}``, written by atomic rename).
  The marker of the generation about to be overwritten is removed first; the new marker is written after the snapshot is
  durable on this rank (``store.flush()``) **and every rank has reached that point** (host barrier) - so a marker for
  round n on any rank implies a complete round-n file on all ranks, and because ranks are never more than one save
  apart, the sets of committed rounds of any two ranks always intersect;
* ``load`` all-gathers the committed rounds and takes the newest round every rank has.
``FLPR_FAULT_EXIT=rank:round[:phase]`` (first attempt of an elastic job only) makes a rank exit hard at ``phase`` =
``round`` (after the round, before the save), ``saving`` (snapshot written, marker not yet) or ``saved`` - the fault
injection behind ``tests/dist_resume_check.py``.
"""
from __future__ import annotations

import json
import os
import random
from typing import Any, Dict

import torch

ACTOR = "_resume"


def _cpu(obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu(v) for v in obj)
    return obj


def client_state(client) -> Dict:
    m = client.model
    op = client.operator
    st = {"train_cnt": client.train_cnt, "test_cnt": client.test_cnt,
          "pipeline": client.task_pipeline.state_dict() if getattr(client, "task_pipeline", None) is not None else None,
          "model": m.full_state(),
          "opt": {"lr": op.optimizer.lr, "step_count": op.optimizer.step_count} if op.optimizer is not None else None,
          "sched": op.scheduler.state_dict() if op.scheduler is not None else None,
          "rng": m.rng.get_state() if getattr(m, "rng", None) is not None else None,
          "model_extra": m.resume_extra() if hasattr(m, "resume_extra") else {},
          "client_extra": client.resume_extra() if hasattr(client, "resume_extra") else {}}
    return st


def load_client_state(client, st: Dict) -> None:
    m = client.model
    op = client.operator
    client.train_cnt, client.test_cnt = int(st["train_cnt"]), int(st["test_cnt"])
    if st.get("pipeline") is not None and getattr(client, "task_pipeline", None) is not None:
        client.task_pipeline.load_state_dict(st["pipeline"])
    m.load_full_state(st["model"])
    if st.get("opt") is not None and op.optimizer is not None:
        op.optimizer.lr, op.optimizer.step_count = float(st["opt"]["lr"]), int(st["opt"]["step_count"])
        op.optimizer.sync_hyper()
    if st.get("sched") is not None and op.scheduler is not None:
        op.scheduler.load_state_dict(st["sched"])
    if st.get("rng") is not None and getattr(m, "rng", None) is not None:
        m.rng.set_state(st["rng"].cpu())
    if hasattr(m, "load_resume_extra"):
        m.load_resume_extra(st.get("model_extra") or {})
    if hasattr(client, "load_resume_extra"):
        client.load_resume_extra(st.get("client_extra") or {})


def save(stage, store, curr_round: int, server, clients, comm) -> None:
    state: Dict[str, Any] = {
        "round": int(curr_round), "world": stage.world, "rank": stage.rank,
        "py_random": random.getstate(), "torch_rng": torch.get_rng_state(),
        "cuda_rng": torch.cuda.get_rng_state(stage.device) if stage.device.type == "cuda" else None,
        "server": {"clients": list(server.clients.keys()),
                   "uploaded": list(getattr(server, "uploaded", [])),
                   "extra": server.resume_extra() if hasattr(server, "resume_extra") else {},
                   "model": server.model.full_state()},
        "clients": {c.client_name: client_state(c) for c in clients},
        "comm": {},
    }
    if comm is not None:
        bufs: Dict[str, Any] = {}
        for name, b in comm.bufs.items():
            if getattr(b, "per_client", False):
                bufs[name] = {int(cid): comm.client_view(name, cid) for cid in comm.local_clients()}
            else:
                bufs[name] = comm.rank_view(name)
        state["comm"] = bufs
    committed = _committed(store, stage.rank)               # {round: generation}
    gen = 1 - committed[max(committed)] if committed else 0  # never the newest committed one: the older / uncommitted
    marker = _marker(store, stage.rank, gen)
    if os.path.exists(marker):
        os.remove(marker)                                   # a torn overwrite must never look committed
    store.save(ACTOR, _gen_name(stage.rank, gen), state, True)
    store.flush()                                           # durable on this rank ...
    maybe_inject_fault(stage, curr_round, "saving")
    _host_barrier(stage)                                    # ... and on every other rank
    tmp = f"{marker}.tmp{os.getpid()}"
    with open(tmp, "w") as f:
        json.dump({"round": int(curr_round), "world": stage.world}, f)
    os.replace(tmp, marker)
    _drop_other_worlds(stage, store)
    maybe_inject_fault(stage, curr_round, "saved")


def _gen_name(rank: int, gen: int) -> str:
    return f"rank{rank}-g{gen}"


def _marker(store, rank: int, gen: int) -> str:
    return store.path(ACTOR, _gen_name(rank, gen)) + ".ok"


def _committed(store, rank: int) -> Dict[int, int]:
    """``{round: generation}`` of this rank's committed manifests (marker present, file present)."""
    out: Dict[int, int] = {}
    for gen in (0, 1):
        marker = _marker(store, rank, gen)
        if os.path.exists(marker) and os.path.exists(store.path(ACTOR, _gen_name(rank, gen))):
            try:
                with open(marker) as f:
                    out[int(json.load(f)["round"])] = gen
            except (OSError, ValueError, KeyError):
                pass
    return out


def _host_barrier(stage) -> None:
    import torch.distributed as dist
    if stage.world > 1 and dist.is_available() and dist.is_initialized():
        dist.barrier()


def maybe_inject_fault(stage, curr_round: int, phase: str) -> None:
    """``FLPR_FAULT_EXIT=rank:round[:phase]``: hard exit of one rank (``os._exit``: no cleanup, no flush - a crash), on
    the first attempt of an elastic job only (``TORCHELASTIC_RESTART_COUNT`` 0 / unset)."""
    spec = os.environ.get("FLPR_FAULT_EXIT")
    if not spec or os.environ.get("TORCHELASTIC_RESTART_COUNT", "0") != "0":
        return
    parts = spec.split(":")
    if int(parts[0]) == stage.rank and int(parts[1]) == int(curr_round) and (parts[2] if len(parts) > 2 else "round") == phase:
        os._exit(17)


def available(store, rank: int) -> bool:
    """A committed manifest of THIS rank exists (rank-local: use :func:`agreed_round` to decide what a job does)."""
    return bool(_committed(store, rank))


def _markers(store):
    """``(path, round, world)`` of every commit marker in the store's ``_resume`` directory."""
    d = os.path.dirname(store.path(ACTOR, "x"))
    out = []
    if os.path.isdir(d):
        for name in os.listdir(d):
            if name.endswith(".ok"):
                try:
                    with open(os.path.join(d, name)) as f:
                        m = json.load(f)
                    out.append((os.path.join(d, name), int(m["round"]), int(m.get("world", 1))))
                except (OSError, ValueError, KeyError):
                    pass
    return out


def written_world(store) -> int:
    """World size of the job that wrote the NEWEST committed manifest in the store's ``_resume`` directory (0: none)."""
    marks = _markers(store)
    return max(marks, key=lambda m: m[1])[2] if marks else 0


def _drop_other_worlds(stage, store) -> None:
    """After the first commit of a re-sharded job: the manifests of the previous world size are history."""
    for path, _, world in _markers(store):
        if world != stage.world:
            for p in (path, path[:-len(".ok")]):
                try:
                    os.remove(p)
                except OSError:
                    pass


def agreed_round(stage, store) -> int:
    """Newest round whose manifest is committed on EVERY rank of the job that wrote it (0: none). Collective when
    ``world > 1``: every rank must call it, and every rank gets the same answer.

    When the world size changed (a GPU was lost for good and the job is restarted on fewer ranks, or grown again) the
    writers' ranks no longer exist: every new rank then reads the markers of ALL old ranks from the shared checkpoint
    directory and the new ranks agree on the minimum of what they see."""
    import torch.distributed as dist
    old_world = written_world(store)
    if old_world in (0, stage.world):
        mine = sorted(_committed(store, stage.rank))
    else:
        per_rank = [set(_committed(store, r)) for r in range(old_world)]
        mine = sorted(set.intersection(*per_rank)) if per_rank else []
    if stage.world > 1 and dist.is_available() and dist.is_initialized():
        everyone = [None] * stage.world
        dist.all_gather_object(everyone, mine)
        common = set(everyone[0]).intersection(*map(set, everyone[1:]))
    else:
        common = set(mine)
    return max(common) if common else 0


def _merged_manifest(stage, store, rnd: int, old_world: int) -> Dict[str, Any]:
    """The manifests of all ``old_world`` writer ranks folded into one picture: clients and per-client symmetric
    buffers are the union (every client was hosted by exactly one old rank), everything replicated (server, rank
    buffers, host RNG streams - identical on every rank by construction) is taken from old rank 0."""
    merged: Dict[str, Any] = {}
    for r in range(old_world):
        st = store.load(ACTOR, _gen_name(r, _committed(store, r)[rnd]))
        assert int(st["round"]) == rnd, (r, st["round"], rnd)
        if r == 0:
            merged = st
            merged["comm"] = dict(st.get("comm") or {})
            continue
        merged["clients"].update(st["clients"])
        for name, val in (st.get("comm") or {}).items():
            if isinstance(val, dict):
                merged["comm"].setdefault(name, {})
                merged["comm"][name] = {**merged["comm"][name], **val}
    merged["world"] = stage.world
    return merged


def load(stage, store, server, clients, comm, rnd: int = None) -> int:
    """Restore the manifest of round ``rnd`` (default: :func:`agreed_round`, the newest one committed on every rank);
    returns that round (0: nothing to resume from - the caller starts at round 1 on the freshly built state). A job
    restarted with a different world size re-shards: each new rank takes the clients it hosts now out of the merged
    manifests of the old ranks (shared checkpoint directory required)."""
    if rnd is None:
        rnd = agreed_round(stage, store)
    if rnd == 0:
        return 0
    old_world = written_world(store)
    if old_world not in (0, stage.world):
        st = _merged_manifest(stage, store, rnd, old_world)
        stage.logger.info(f"Resume: re-sharding the round-{rnd} manifests of {old_world} ranks over {stage.world}.")
    else:
        st = store.load(ACTOR, _gen_name(stage.rank, _committed(store, stage.rank)[rnd]))
    assert int(st["round"]) == rnd, (st["round"], rnd)
    random.setstate(st["py_random"])
    torch.set_rng_state(st["torch_rng"])
    if st.get("cuda_rng") is not None and stage.device.type == "cuda":
        torch.cuda.set_rng_state(st["cuda_rng"], stage.device)
    sv = st["server"]
    for name in sv["clients"]:
        if name not in server.clients:
            server.clients[name] = {"resumed": True}
    if hasattr(server, "uploaded"):
        server.uploaded = list(sv["uploaded"])
    server.model.load_full_state(sv["model"])
    if hasattr(server, "load_resume_extra"):
        server.load_resume_extra(sv.get("extra") or {})
    by_name = {c.client_name: c for c in clients}
    for name, cs in st["clients"].items():
        if name in by_name:
            load_client_state(by_name[name], cs)
    if comm is not None:
        for name, val in (st.get("comm") or {}).items():
            if name not in comm.bufs:
                continue
            if isinstance(val, dict):
                for cid, t in val.items():
                    if comm.owner(int(cid)) != comm.rank:
                        continue                                  # (re-sharded: this slot lives on another rank now)
                    comm.client_view(name, int(cid)).copy_(t.to(comm.client_view(name, int(cid)).device))
            else:
                comm.rank_view(name).copy_(val.to(comm.rank_view(name).device))
        comm.barrier()
    return int(st["round"])
