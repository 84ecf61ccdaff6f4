"""Ceiling check for the checkpoint writers: aggregate write bandwidth into the RAM disk with N processes."""
import multiprocessing as mp, os, sys, time
import numpy as np


def work(i, nbytes, reps, q):
    buf = np.ones(nbytes, dtype=np.uint8)
    t0 = time.perf_counter()
    for r in range(reps):
        p = f"/dev/shm/flpr_bw_{i}_{r}"
        fd = os.open(p, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        off = 0
        mv = memoryview(buf)
        while off < nbytes:
            off += os.write(fd, mv[off:off + (256 << 20)])
        os.close(fd)
    dt = time.perf_counter() - t0
    for r in range(reps):
        os.remove(f"/dev/shm/flpr_bw_{i}_{r}")
    q.put(dt)


if __name__ == "__main__":
    print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
    os.system("df -h /dev/shm | tail -1; free -g | head -2")
    for n in (1, 4, 12, 24):
        q = mp.Queue()
        ps = [mp.Process(target=work, args=(i, 512 << 20, 2, q)) for i in range(n)]
        t0 = time.perf_counter()
        [p.start() for p in ps]
        dts = [q.get() for _ in ps]
        [p.join() for p in ps]
        wall = time.perf_counter() - t0
        print(f"{n:2d} writers: {n * 2 * 0.5 / max(dts):6.1f} GiB/s aggregate (slowest writer {max(dts):.2f}s, wall {wall:.2f}s)")
