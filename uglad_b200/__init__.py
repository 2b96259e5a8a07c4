"""uglad_b200: B200-native (sm_100a) implementation of uGLAD's unrolled GLAD hot path.

Drop-in surface (same names and argument meaning as the reference's uglad package):
    uglad_b200.main.uGLAD_GL / uGLAD_multitask / run_uGLAD_* / forward_uGLAD / loss_uGLAD
    uglad_b200.glad.glad.glad / get_optimizers,  uglad_b200.glad.glad_params.GladParams
    uglad_b200.utils.prepare_data.get_covariance / convert_to_torch / process_table
Everything numeric runs in libuglad_b200.so (include/uglad_b200.h); there is no CPU path.
"""
__version__ = "0.1.0"
