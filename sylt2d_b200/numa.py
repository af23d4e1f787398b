"""Host-side placement helper for the end-to-end path: a rank that streams body states through pinned host buffers every
step (sylt_batch_step_host6) should run, and allocate those buffers, on the NUMA node its GPU hangs off — otherwise every
byte crosses the inter-socket link that all GPUs of the other socket share.  Pure host logic (sysfs + nvidia-smi); no CUDA."""
import os
import subprocess


def _parse_cpulist(s):
    cpus = set()
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_info(index):
    """{'bus_id', 'numa_node', 'cpus'} of GPU `index` (as numbered by nvidia-smi / CUDA_VISIBLE_DEVICES order), or {} if unknown."""
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                phys = int(ids[index])
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(phys)],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
        if not out:
            return {}
        bus = out[0].strip().lower()          # 00000000:1B:00.0 -> sysfs uses a 4-digit domain
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        return {"bus_id": bus, "numa_node": node, "cpus": sorted(cpus)}
    except Exception:
        return {}


def bind_to_gpu_numa(index):
    """Restrict this process to the CPUs local to GPU `index` (memory then comes from that node: first touch, local policy).
    Returns the info dict, with 'bound': True when the affinity was changed."""
    info = gpu_numa_info(index)
    cpus = set(info.get("cpus", [])) & os.sched_getaffinity(0) if info else set()
    info = dict(info)
    info["bound"] = False
    if cpus and info.get("numa_node", -1) >= 0:
        try:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
        except OSError:
            pass
    info.pop("cpus", None)
    info["n_cpus"] = len(cpus)
    return info
