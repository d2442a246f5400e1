"""geneface_b200 -- B200-native (sm_100a) volumetric talking-head renderer behind GeneFace's RAD-NeRF API.

Public surface mirrors the reference modules on the hot path:
    geneface_b200.raymarching   <- modules/radnerfs/raymarching/raymarching.py
    geneface_b200.encoders      <- modules/radnerfs/encoders/*
    geneface_b200.renderer      <- modules/radnerfs/{renderer,radnerf,radnerf_torso}.py
    geneface_b200.cond_encoder  <- modules/radnerfs/cond_encoder.py
    geneface_b200.utils         <- the functions of modules/radnerfs/utils.py the path uses
    geneface_b200.sequence      <- inference/nerfs/base_nerf_infer.py frame sharding (multi-GPU)
All compute goes through libgfrender.so (include/gfrender.h); there is no CPU or eager fallback.
"""
__version__ = "0.1.0"
