"""stpy_b200: the Gaussian-process hot path of stpy on NVIDIA B200 (sm_100a).

Import paths mirror the reference package:
    stpy.kernels.KernelFunction                                   -> stpy_b200.kernels.KernelFunction
    stpy.continuous_processes.gauss_procc.GaussianProcess         -> stpy_b200.continuous_processes.gauss_procc.GaussianProcess
    stpy.embeddings.embedding.RFFEmbedding                        -> stpy_b200.embeddings.embedding.RFFEmbedding
    stpy.continuous_processes.kernelized_features.KernelizedFeatures
                                                                  -> stpy_b200.continuous_processes.kernelized_features.KernelizedFeatures
    stpy.continuous_processes.categorical_mixture.CategoricalMixture
                                                                  -> stpy_b200.continuous_processes.categorical_mixture.CategoricalMixture
    stpy.continuous_processes.mkl_estimator.MultipleKernelLearner -> stpy_b200.continuous_processes.mkl_estimator.MultipleKernelLearner
    stpy.continuous_processes.nystrom_fea.NystromFeatures         -> stpy_b200.continuous_processes.nystrom_fea.NystromFeatures
    stpy.kernel_functions.{squared_exponential_kernel,ard_kernel} -> stpy_b200.kernel_functions.{...}
"""
__version__ = "0.2.0"
