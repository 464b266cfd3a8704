import os, sys, time, glob, subprocess, torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
print("nodes", nodes, "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
print("bus", bus)
for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    dev = p.split("/")[-2]
    try:
        cls = open(p.replace("numa_node", "class")).read().strip()
        if cls.startswith("0x0302") or cls.startswith("0x0300"):
            print(dev, cls, open(p).read().strip(), open(p.replace("numa_node", "local_cpulist")).read().strip())
    except Exception as e:
        pass
def bw():
    big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); big.fill_(1)
    dbig = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): dbig.copy_(big, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): dbig.copy_(big, non_blocking=True)
    torch.cuda.synchronize(); return (64 << 20) * 10 / (time.perf_counter() - t0) / 1e9
print("default h2d GB/s", bw())
for nd in nodes:
    cl = open(nd + "/cpulist").read().strip()
    cpus = set()
    for part in cl.split(","):
        a, _, b = part.partition("-"); cpus |= set(range(int(a), int(b or a) + 1))
    try:
        os.sched_setaffinity(0, cpus)
        print(nd, cl, "h2d GB/s", bw())
    except Exception as e:
        print(nd, "affinity failed", e)
