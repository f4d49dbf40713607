"""Layout of libaom's FRAME_CONTEXT (av1/common/entropymode.h order), used to slice the buffer
filled by av1_init_mode_probs() into named default CDF arrays.  (name, dims, nsyms)"""
COEF = [
    ("txb_skip", (5, 13), 2), ("eob_extra", (5, 2, 9), 2), ("dc_sign", (2, 3), 2),
    ("eob_pt_16", (2, 2), 5), ("eob_pt_32", (2, 2), 6), ("eob_pt_64", (2, 2), 7), ("eob_pt_128", (2, 2), 8),
    ("eob_pt_256", (2, 2), 9), ("eob_pt_512", (2, 2), 10), ("eob_pt_1024", (2, 2), 11),
    ("coeff_base_eob", (5, 2, 4), 3), ("coeff_base", (5, 2, 42), 4), ("coeff_br", (5, 2, 21), 4),
]
NMV = [("joints", (), 4)]
for c in range(2):
    NMV += [(f"c{c}_classes", (), 11), (f"c{c}_class0_fp", (2,), 4), (f"c{c}_fp", (), 4), (f"c{c}_sign", (), 2),
            (f"c{c}_class0_hp", (), 2), (f"c{c}_hp", (), 2), (f"c{c}_class0", (), 2), (f"c{c}_bits", (10,), 2)]
MODE = [
    ("newmv", (6,), 2), ("zeromv", (2,), 2), ("refmv", (6,), 2), ("drl", (3,), 2),
    ("inter_compound_mode", (8,), 8), ("compound_type", (22,), 2), ("wedge_idx", (22,), 16),
    ("interintra", (4,), 2), ("wedge_interintra", (22,), 2), ("interintra_mode", (4,), 4),
    ("motion_mode", (22,), 3), ("obmc", (22,), 2),
    ("palette_y_size", (7,), 7), ("palette_uv_size", (7,), 7),
    ("palette_y_color_index", (7, 5), 8), ("palette_uv_color_index", (7, 5), 8),
    ("palette_y_mode", (7, 3), 2), ("palette_uv_mode", (2,), 2),
    ("comp_inter", (5,), 2), ("single_ref", (3, 6), 2), ("comp_ref_type", (5,), 2),
    ("uni_comp_ref", (3, 3), 2), ("comp_ref", (3, 3), 2), ("comp_bwdref", (3, 2), 2),
    ("txfm_partition", (21,), 2), ("compound_index", (6,), 2), ("comp_group_idx", (6,), 2),
    ("skip_mode", (3,), 2), ("skip", (3,), 2), ("intra_inter", (4,), 2),
] + [("mv_" + n, d, s) for n, d, s in NMV] + [("dv_" + n, d, s) for n, d, s in NMV] + [
    ("intrabc", (), 2),
    ("seg_pred", (3,), 2), ("seg_spatial", (3,), 8),
    ("filter_intra", (22,), 2), ("filter_intra_mode", (), 5),
    ("switchable_restore", (), 3), ("wiener_restore", (), 2), ("sgrproj_restore", (), 2),
    ("y_mode", (4,), 13), ("uv_mode", (2, 13), 14), ("partition", (20,), 10),
    ("switchable_interp", (16,), 3), ("kf_y_mode", (5, 5), 13), ("angle_delta", (8,), 7),
    ("tx_size", (4, 3), 3), ("delta_q", (), 4), ("delta_lf_multi", (4,), 4), ("delta_lf", (), 4),
    ("intra_ext_tx", (3, 4, 13), 16), ("inter_ext_tx", (4, 4), 16),
    ("cfl_sign", (), 8), ("cfl_alpha", (6,), 16),
]


def count(dims):
    n = 1
    for d in dims:
        n *= d
    return n


def total(layout):
    return sum(count(d) * (s + 1) for _, d, s in layout)
