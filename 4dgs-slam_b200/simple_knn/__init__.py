"""B200-native drop-in for the reference's ``simple_knn`` extension (submodules/simple-knn): ``from simple_knn._C import distCUDA2``
keeps working unchanged (gaussian_splatting/scene/gaussian_model.py:18).  The compute lives in libg4r.so (csrc/knn.cu,
C ABI ``g4r_knn_mean_dist2``); there is no CPU / PyTorch fallback."""
