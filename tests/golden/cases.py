"""The golden stream set: generator parameters only (streams are regenerated deterministically)."""
CASES = {
    "cif420_ipb": (352, 288, 1, dict(seed=101, n_gops=2, gop_n=9, gop_m=3)),
    "cif422_ipb": (352, 288, 2, dict(seed=102, n_gops=2, gop_n=9, gop_m=3)),
    "cif444_ipb": (352, 288, 3, dict(seed=103, n_gops=2, gop_n=9, gop_m=3)),
    "qcif420_stress": (176, 144, 1, dict(seed=104, qscale_code_max=31, pct_big_levels=40, gop_n=12, gop_m=3)),
    "qcif444_altscan_nonlinear": (176, 144, 3, dict(seed=105, alternate_scan=1, q_scale_type=1, intra_dc_precision=3, qscale_code_max=31, gop_n=9, gop_m=3)),
    "sd420_intra": (720, 480, 1, dict(seed=106, intra_only=1, gop_n=4)),
    "sd422_longmv": (720, 480, 2, dict(seed=107, gop_n=7, gop_m=3, mv_range=120, pct_skipped=30)),
    "tiny420_m1": (48, 32, 1, dict(seed=108, gop_n=10, gop_m=1)),
    "natural420": (640, 368, 1, dict(seed=109, mode=1, n_gops=2, gop_n=15, gop_m=3)),
    # texture mode: a translating procedural texture + noise, really encoded (forward DCT, quantiser_scale 4..8)
    "texture420": (640, 368, 1, dict(seed=113, mode=2, n_gops=2, gop_n=9, gop_m=3, pct_intra_in_pb=3)),
    "texture444": (352, 288, 3, dict(seed=114, mode=2, n_gops=1, gop_n=7, gop_m=3, q_scale_type=1, alternate_scan=1)),
    "cif420_userdata": (352, 288, 1, dict(seed=115, n_gops=2, gop_n=6, gop_m=3, user_data_bytes=37)),
    # field DCT: frame pictures with frame_pred_frame_dct = 0, frame-based prediction, dct_type = 1 in about half of the
    # coded macroblocks.  The reference decodes these for 4:2:0 and 4:2:2 (mb_decoder.cpp:172-189); its 4:4:4 path writes
    # outside the macroblock (:194-195, heap corruption at the bottom row), so there is no 4:4:4 entry here
    "fielddct420_ipb": (352, 288, 1, dict(seed=116, n_gops=2, gop_n=9, gop_m=3, pct_field_dct=50)),
    "fielddct422_altscan": (352, 288, 2, dict(seed=117, n_gops=2, gop_n=9, gop_m=3, pct_field_dct=60, alternate_scan=1, q_scale_type=1)),
    "fielddct420_texture": (640, 368, 1, dict(seed=118, mode=2, n_gops=2, gop_n=9, gop_m=3, pct_intra_in_pb=3, pct_field_dct=40)),
    "tall420_vpos_ext": (32, 2816, 1, dict(seed=112, gop_n=4, gop_m=3)),
    "hd420_ipb": (1920, 1088, 1, dict(seed=110, gop_n=7, gop_m=3)),
    "hd422_ipb": (1920, 1088, 2, dict(seed=111, gop_n=4, gop_m=3)),
}
