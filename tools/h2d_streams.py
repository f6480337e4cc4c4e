import torch, time
dev = torch.device("cuda", 0)
imgs = [torch.randn(32, 688, 464).pin_memory() for _ in range(5)]
dst = torch.empty(5, 32, 688, 464, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]
for ns in (1, 2, 4):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(5):
            with torch.cuda.stream(streams[i % ns]):
                dst[i].copy_(imgs[i], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("streams %d: %.2f ms for 204 MB = %.1f GB/s" % (ns, dt * 1e3, 5 * 32 * 688 * 464 * 4 / dt / 1e9))
# split each image in 2 halves over 2 streams
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(5):
        for hlf in range(2):
            with torch.cuda.stream(streams[hlf]):
                dst[i, hlf * 16:(hlf + 1) * 16].copy_(imgs[i][hlf * 16:(hlf + 1) * 16], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("halves on 2 streams: %.2f ms = %.1f GB/s" % (dt * 1e3, 5 * 32 * 688 * 464 * 4 / dt / 1e9))
