"""Inputs of the host-logic golden cases (shared by make_golden_host.py and tests/test_host_logic.py).
Each case: (name, icp_type, factory) where factory() -> (source, target, T_init, weight)."""
import torch


def _rand(gen, *shape, dtype=torch.float32):
    return torch.rand(shape, generator=gen, dtype=dtype) * 10.0 - 5.0


def host_cases():
    cases = []

    def add(name, icp_type, fn):
        cases.append((name, icp_type, fn))

    def ragged_lists(cols_t, dtype, with_w, with_T=True):
        def make():
            g = torch.Generator().manual_seed(11)
            src = [_rand(g, 5, 3, dtype=dtype), _rand(g, 9, 3, dtype=dtype), _rand(g, 2, 3, dtype=dtype)]
            tgt = [_rand(g, 7, cols_t, dtype=dtype), _rand(g, 4, cols_t, dtype=dtype), _rand(g, 11, cols_t, dtype=dtype)]
            T = [torch.eye(4, dtype=dtype) for _ in range(3)] if with_T else None
            w = [torch.rand(5, generator=g, dtype=dtype), torch.rand(9, generator=g, dtype=dtype),
                 torch.rand(2, generator=g, dtype=dtype)] if with_w else None
            return src, tgt, T, w
        return make
    add("ragged_pt2pl_f32", "pt2pl", ragged_lists(6, torch.float32, False))
    add("ragged_pt2pl_f64_weights", "pt2pl", ragged_lists(6, torch.float64, True))
    add("ragged_pt2pt_f32_weights", "pt2pt", ragged_lists(3, torch.float32, True))
    add("ragged_pt2pt_6col_target", "pt2pt", ragged_lists(6, torch.float32, False))
    add("ragged_no_T", "pt2pl", ragged_lists(6, torch.float32, False, with_T=False))

    def single_pair(cols_t, with_w):
        def make():
            g = torch.Generator().manual_seed(12)
            return (_rand(g, 6, 3, dtype=torch.float64), _rand(g, 8, cols_t, dtype=torch.float64), torch.eye(4, dtype=torch.float64),
                    torch.rand(6, generator=g, dtype=torch.float64) if with_w else None)
        return make
    add("single_pair_pt2pl", "pt2pl", single_pair(6, False))
    add("single_pair_pt2pt_weights", "pt2pt", single_pair(3, True))

    def batched(cols_s, cols_t, with_w):
        def make():
            g = torch.Generator().manual_seed(13)
            return (_rand(g, 2, 6, cols_s), _rand(g, 2, 5, cols_t), torch.eye(4).repeat(2, 1, 1),
                    torch.rand((2, 6), generator=g) if with_w else None)
        return make
    add("batched_pt2pl", "pt2pl", batched(3, 6, False))
    add("batched_pt2pl_weights", "pt2pl", batched(3, 6, True))
    add("batched_pt2pt_weights", "pt2pt", batched(3, 3, True))
    add("batched_source_with_normals", "pt2pl", batched(6, 6, False))

    def with_empty(which):
        def make():
            g = torch.Generator().manual_seed(14)
            src = [_rand(g, 4, 3), _rand(g, 6, 3)]
            tgt = [_rand(g, 5, 6), _rand(g, 3, 6)]
            if which == "source":
                src[0] = torch.empty((0, 3))
            elif which == "target":
                tgt[1] = torch.empty((0, 6))
            else:
                src[1] = torch.empty((0, 3))
                tgt[1] = torch.empty((0, 6))
            return src, tgt, [torch.eye(4), torch.eye(4)], None
        return make
    add("empty_source_cloud", "pt2pl", with_empty("source"))
    add("empty_target_cloud", "pt2pl", with_empty("target"))
    add("empty_both", "pt2pl", with_empty("both"))

    def zero_rows():
        g = torch.Generator().manual_seed(15)
        s = _rand(g, 2, 6, 3)
        s[0, 4:] = 0.0                             # zero rows of a dense batch: padding under source_zeroes_are_pad
        return s, _rand(g, 2, 5, 6), torch.eye(4).repeat(2, 1, 1), None
    add("dense_batch_with_zero_rows", "pt2pl", zero_rows)

    def negative_sources():
        g = torch.Generator().manual_seed(16)
        src = [-torch.rand((4, 3), generator=g) - 1.0, -torch.rand((6, 3), generator=g) - 1.0]   # max(source) < 0: the pad-value hazard
        tgt = [_rand(g, 5, 6), _rand(g, 3, 6)]
        return src, tgt, [torch.eye(4), torch.eye(4)], None
    add("negative_sources_pad_value", "pt2pl", negative_sources)

    def weight_length_mismatch():
        g = torch.Generator().manual_seed(17)
        return [_rand(g, 4, 3), _rand(g, 6, 3)], [_rand(g, 5, 6), _rand(g, 3, 6)], None, [torch.rand(4, generator=g)]
    add("weight_list_length_mismatch", "pt2pl", weight_length_mismatch)
    return cases
