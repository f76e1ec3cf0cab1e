"""SM partitions (include/dots_ocr_b200.h "SM partitions", csrc/partition.cu): kernels launched into a partition's stream give the
results of the whole-device launch, a decode-style graph can be captured on a partition stream and replayed there, and both
partitions really run at the same time."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def parts():
    from dots_ocr_b200 import ops
    sp, sd, n_p, n_d = ops.partition(96)          # process-lifetime (ops.partition): the pipeline tests reuse the same split
    yield sp, sd, n_p, n_d
    torch.cuda.synchronize()


def test_partition_sizes_and_results(parts):
    from dots_ocr_b200 import ops
    sp, sd, n_p, n_d = parts
    total = torch.cuda.get_device_properties(0).multi_processor_count
    assert n_p == 96 and 8 <= n_d <= total - 96
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn((4096, 1536), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((4608, 1536), device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    ref = ops.gemm(a, w)
    torch.cuda.synchronize()
    for st, n in ((sp, n_p), (sd, n_d)):
        with ops.on_partition(st, n):
            got = ops.gemm(a, w)
        torch.cuda.synchronize()
        assert torch.equal(ref, got)


def test_graph_on_a_partition_stream(parts):
    from dots_ocr_b200 import ops
    sp, sd, n_p, n_d = parts
    x = (torch.randn((64, 1536), device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn((2048, 1536), device="cuda") * 0.03).to(torch.bfloat16)
    ref = ops.gemm_skinny(x, w, 1, out_bf16=torch.empty((64, 2048), device="cuda", dtype=torch.bfloat16)).clone()
    out = torch.zeros((64, 2048), device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()
    with ops.on_partition(sd, n_d):
        gr = ops.capture(lambda: ops.gemm_skinny(x, w, 1, out_bf16=out))
        gr.launch()
    torch.cuda.synchronize()
    assert torch.equal(ref, out)
    del gr


def test_partitions_run_concurrently(parts):
    """Two long GEMM loops, one per partition: together they must take clearly less than the sum of each alone."""
    from dots_ocr_b200 import ops
    sp, sd, n_p, n_d = parts
    a = torch.randn((16384, 1536), device="cuda").to(torch.bfloat16)
    w = (torch.randn((4608, 1536), device="cuda") * 0.03).to(torch.bfloat16)
    outs = [torch.empty((16384, 4608), device="cuda", dtype=torch.bfloat16) for _ in range(2)]

    def loop(st, n, out, reps):
        with ops.on_partition(st, n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ops.gemm(a, w, out=out)
            e1.record()
        return e0, e1

    for st, n, o in ((sp, n_p, outs[0]), (sd, n_d, outs[1])):
        loop(st, n, o, 2)
    torch.cuda.synchronize()
    e = loop(sp, n_p, outs[0], 40); torch.cuda.synchronize(); t_p = e[0].elapsed_time(e[1])
    e = loop(sd, n_d, outs[1], 40); torch.cuda.synchronize(); t_d = e[0].elapsed_time(e[1])
    ea = loop(sp, n_p, outs[0], 40)
    eb = loop(sd, n_d, outs[1], 40)
    torch.cuda.synchronize()
    both = max(ea[0].elapsed_time(eb[1]), ea[0].elapsed_time(ea[1]))
    print(f"alone {t_p:.2f} + {t_d:.2f} ms, together {both:.2f} ms")
    # serialised execution would give the sum (ratio 1.0); measured 0.78 (the board's power cap slows both loops when they overlap)
    assert both < 0.92 * (t_p + t_d)


def test_partition_is_cached_and_a_second_split_is_refused(parts):
    from dots_ocr_b200 import ops
    again = ops.partition(96)
    assert again[0] is parts[0] and again[2:] == parts[2:]
    with pytest.raises(RuntimeError):
        ops.partition(64)
