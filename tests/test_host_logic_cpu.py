"""Host-side logic that needs no GPU: the colour-MLP first-layer composition / gradient split, the backward plan of the
rows MLP, compositing run-length dispatch invariants (through the shared library's pure host helpers where exported)."""
import numpy as np
import torch

from simple_rf_b200.nerf_program import PackedRowsMLP, DgradProgram, MlpProgram
from simple_rf_b200 import tile_images


def _lin(o, i, g):
    return (torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) / i ** 0.5


def test_composed_first_layer_and_gradient_split_match_autograd():
    """W0' = [W0[:, :F] B | W0[:, F:]] (reference: basis_matrix_color then mlp.0, SimpleTensoRF09.py:1263, :1389-1393);
    split_first_layer_grad must be the exact chain rule of that composition."""
    g = torch.Generator().manual_seed(0)
    m = PackedRowsMLP(72, 27, 3, prefix='mlp')
    w0 = _lin(128, 30, g).requires_grad_()
    basis = _lin(27, 72, g).requires_grad_()
    x = torch.rand(50, 75, generator=g, dtype=torch.float64)
    # two-step reference: features = prod @ B^T ; h = [features | vd] @ W0^T
    ref = torch.cat([x[:, :72] @ basis.t(), x[:, 72:]], 1) @ w0.t()
    comp = x @ m.composed_first_layer(w0, basis).t()
    assert torch.allclose(ref, comp, atol=1e-12)
    up = torch.rand(ref.shape, generator=g, dtype=torch.float64)
    gw0_ref, gb_ref = torch.autograd.grad((ref * up).sum(), [w0, basis])
    g_comp = up.t() @ x                                        # d loss / d W0'
    flat = torch.zeros(m.flat_size, dtype=torch.float64)
    flat[:128 * 75] = g_comp.reshape(-1)
    gw0, gb = m.split_first_layer_grad(flat, w0.detach(), basis.detach())
    assert torch.allclose(gw0, gw0_ref, atol=1e-10) and torch.allclose(gb, gb_ref, atol=1e-10)


def test_rows_mlp_programs_are_consistent():
    """Forward program / backward plan of the colour MLP: slots, widths and image counts line up with what
    srf_mlp_rows_fwd saves and srf_nerf_mlp_dgrad / _wgrad read (validated again by the library at launch)."""
    m = PackedRowsMLP(72, 27, 3, prefix='mlp')
    p: MlpProgram = m.program
    assert p.num_layers == 2 and p.views_degree == -2
    assert [p.layers[i].n for i in range(2)] == [128, 128]
    assert list(p.layers[0].kblock_region[:2]) == [0, 5] and list(p.layers[0].kblock_ksteps[:2]) == [4, 1]
    assert (p.layers[0].save_slot, p.layers[1].save_slot, m.e_slot, m.v_slot, m.act_slots) == (1, 3, 0, 5, 6)
    d: DgradProgram = m.backward_plan.program
    assert d.num_layers == 2 and d.top_width == 128 and d.head_kind == 1 and d.top_mask_slot == 3
    assert (d.layers[0].n_out, d.layers[0].mask_slot, d.layers[0].dz_slot) == (128, 1, 4)
    assert (d.layers[1].n_out, d.layers[1].mask_slot, d.layers[1].dz_slot, d.layers[1].rows_cols) == (128, -1, -1, 72)
    items = m.backward_plan.items
    covered = np.zeros((128, 75), dtype=int)
    for it in items:
        if it.dw_offset == m.offs['mlp.0.weight']:
            covered[:it.out_rows, it.w_col0:it.w_col0 + it.in_cols] += 1
    assert (covered == 1).all()                                # every element of dW0' written by exactly one work item
    # transposed-weight gather of the last backward layer: rows >= 75 of W0'^T images read the zero element
    gt = m.backward_plan.gather_t.reshape(4, 128, 64)
    zero = m.flat_size
    dec = tile_images.decode if hasattr(tile_images, 'decode') else None
    assert (gt[2:] == zero).any() and not (gt[:2] == zero).any()


def test_one_block_rows_program():
    m = PackedRowsMLP(24, 27, 3, prefix='mlp')                 # 27 input columns: a single K block, no region 5
    assert m.program.views_degree == -1 and m.v_slot == -1 and m.act_slots == 5
    assert len(m.backward_plan.items) == 3


def test_input_gradient_sources_cover_every_encoding_consumer(golden_configs):
    """Learnable cameras: the table `srf_nerf_mlp_input_grad` walks (nerf_program.build_backward_plan) names, for every layer that reads
    an encoding image, the dZ images of that layer and the weight column every image column multiplies — checked against the layer
    shapes of the three shipped MLP variants (SimpleNeRF17.py:616-667; the sigma-encoding variant routes the upper octaves to the view layer)."""
    from simple_rf_b200.nerf_program import PackedMLP
    configs, _ = golden_configs('nerf')
    m = configs['model']
    variants = [m['coarse_model'], m['augmentations'][0]['coarse_model'], m['augmentations'][1]['coarse_model']]
    seen_split = False
    for cfg in variants:
        packed = PackedMLP(cfg)
        plan = packed.backward_plan
        full = 3 * (2 * cfg['points_positional_encoding_degree'] + 1)
        pts_in = 3 * (2 * cfg['points_sigma_positional_encoding_degree'] + 1) if 'points_sigma_positional_encoding_degree' in cfg else full
        extra = full - pts_in
        seen_split = seen_split or extra > 0
        by_weight = {}
        for s in plan.input_sources:
            name = next(n for n, o in packed.offs.items() if o == s.w_offset and n.endswith('.weight'))
            assert s.in_total == packed.shapes[name][1] and s.dz_images * 64 == packed.shapes[name][0]
            assert 0 <= s.dz_slot and s.dz_slot + s.dz_images <= plan.dz_slots
            cols = list(s.cols)
            if s.target == 1:
                assert all(c < 0 for c in cols[32:]), 'the view image holds 32 columns'
            by_weight.setdefault(name, []).append((s.target, cols))
        want = {'pts_linears.0.weight': [(0, list(range(pts_in)) + [-1] * (64 - pts_in))]}
        skip_layer = 'pts_linears.5.weight'
        if packed.shapes.get(skip_layer, (0, 0))[1] == 256 + pts_in:
            want[skip_layer] = [(0, list(range(pts_in)) + [-1] * (64 - pts_in))]          # cat([input_pts, h]): encoding first
        if packed.use_views:
            v = 3 * (2 * cfg['views_positional_encoding_degree'] + 1)
            entries = []
            if extra > 0:
                entries.append((0, [-1] * pts_in + [256 + i for i in range(extra)] + [-1] * (64 - full)))
            entries.append((1, [256 + extra + j for j in range(v)] + [-1] * (64 - v)))
            want['views_linears.0.weight'] = entries
        assert by_weight == want, (cfg, by_weight.keys())
    assert seen_split
