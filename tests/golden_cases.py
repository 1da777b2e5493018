"""Table of golden cases: name -> model specification in reference vocabulary.

The same table drives (a) tests/golden/make_golden.py, which instantiates the *reference's* classes
of these names under the jax shim, (b) the oracle builders and (c) the product (CUDA) builders in the
tests — the class names and keyword arguments are identical in all three.
A prior spec is (class name, kwargs) or ("Sum", [spec, ...]).
"""

AIR_M52 = ("Matern52", dict(variance=2., lengthscale=5.5e4))
AIR_QP365 = ("QuasiPeriodicMatern32", dict(variance=1., lengthscale_periodic=2., period=365., lengthscale_matern=1.5e4))
AIR_QP7 = ("QuasiPeriodicMatern32", dict(variance=1., lengthscale_periodic=2., period=7., lengthscale_matern=30 * 365.))
M32_11 = ("Matern32", dict(variance=1.0, lengthscale=1.0))
M52_11 = ("Matern52", dict(variance=1.0, lengthscale=1.0))
M52_15 = ("Matern52", dict(variance=1.0, lengthscale=5.0))
M32_154 = ("Matern32", dict(variance=1.5, lengthscale=4.0))
ST52 = ("SpatioTemporalMatern52", dict(variance=0.3, lengthscale_time=0.3, lengthscale_space=0.3))
ST52_M25 = ("SpatioTemporalMatern52", dict(variance=0.3, lengthscale_time=0.3, lengthscale_space=0.3,
                                            z=[-3.0 + 0.25 * i for i in range(25)]))
M12_12 = ("Matern12", dict(variance=1.0, lengthscale=2.0))
M72_15 = ("Matern72", dict(variance=1.0, lengthscale=5.0))
COS_05 = ("Cosine", dict(frequency=[0.5]))   # the reference indexes hyp[0] (priors.py:382): frequency must be array-like
PER_3 = ("Periodic", dict(variance=0.8, lengthscale=1.5, period=3.0))
QPM12 = ("QuasiPeriodicMatern12", dict(variance=1.0, lengthscale_periodic=1.5, period=20., lengthscale_matern=40.))
SUB12 = ("SubbandMatern12", dict(variance=0.7, lengthscale=3.0, radial_frequency=1.3))
SUB32 = ("SubbandMatern32", dict(variance=1.1, lengthscale=4.0, radial_frequency=0.4))
SUB52 = ("SubbandMatern52", dict(variance=0.9, lengthscale=3.0, radial_frequency=0.7))
SP52 = ("SpatialMatern52", dict(variance=0.4, lengthscale=0.6))
SP32 = ("SpatialMatern32", dict(variance=0.4, lengthscale=0.8, z=[-3.0 + 0.5 * i for i in range(13)]))
M52FV = ("Matern52FixedVar", dict(variance=1.2, lengthscale=4.0))
SUBFV = ("SubbandExponentialFixedVar", dict(variance=0.6, lengthscale=3.0, radial_frequency=1.1))
M32_31 = ("Matern32", dict(variance=3.0, lengthscale=1.0))
M52_205 = ("Matern52", dict(variance=2.0, lengthscale=0.5))
HETERO = ("HeteroscedasticNoise", dict())
POISSON = ("Poisson", dict())
PROBIT = ("Probit", dict())
LOGIT = ("Bernoulli", dict(link="logit"))

# name: (data set, prior, likelihood, (method class, kwargs), iterations, kat)
CASES = {
    "k1_regression_m52_gauss_pep": ("regression", M52_15, ("Gaussian", dict(variance=0.5)), ("EP", dict(power=0.5)), 2, 797.76),
    "k2_coal200_m52_poisson_eks": ("coal200", M52_11, POISSON, ("EKS", dict()), 3, 295.62),
    "k3_classif_m52_logit_pep_ut": ("classification", M52_15, LOGIT, ("ExpectationPropagation", dict(power=0.9, intmethod="UT")), 2, 469.46),
    "c2_coal333_m32_poisson_pep_gh": ("coal333", M32_11, POISSON, ("EP", dict(power=0.5, intmethod="GH")), 3, None),
    "coal333_m52_poisson_ep1_ut": ("coal333", M52_11, POISSON, ("EP", dict(power=1.0, intmethod="UT")), 3, None),
    "coal333_m52_poisson_ep001_gh_damped": ("coal333", M52_11, POISSON, ("EP", dict(power=0.01, damping=0.5, intmethod="GH")), 3, None),
    "coal333_m52_poisson_eep1": ("coal333", M52_11, POISSON, ("EEP", dict(power=1)), 3, None),
    "coal333_m52_poisson_eep05_damped": ("coal333", M52_11, POISSON, ("EEP", dict(power=0.5, damping=0.7)), 3, None),
    "coal333_m32_poisson_eks_damped": ("coal333", M32_11, POISSON, ("EKS", dict(damping=0.5)), 3, None),
    "coal333_m52_poisson_ghep1": ("coal333", M52_11, POISSON, ("GHEP", dict(power=1)), 3, None),
    "coal333_m52_poisson_ghep05_damped": ("coal333", M52_11, POISSON, ("GHEP", dict(power=0.5, damping=0.6)), 3, None),
    "coal333_m52_poisson_ghks": ("coal333", M52_11, POISSON, ("GHKS", dict()), 3, None),
    "coal333_m32_poisson_uep05": ("coal333", M32_11, POISSON, ("UEP", dict(power=0.5)), 3, None),
    "coal333_m52_poisson_uks": ("coal333", M52_11, POISSON, ("UKS", dict()), 3, None),
    "coal333_m52_poisson_vi_gh": ("coal333", M52_11, POISSON, ("VI", dict(intmethod="GH")), 3, None),
    "coal333_m32_poisson_vi_ut_damped": ("coal333", M32_11, POISSON, ("VI", dict(intmethod="UT", damping=0.5)), 3, None),
    "probit300_m52_ep1": ("classification300", M52_15, PROBIT, ("EP", dict(power=1.0)), 3, None),
    "probit300_m52_pep05_gh": ("classification300", M52_15, PROBIT, ("EP", dict(power=0.5)), 3, None),
    "probit300_m52_vi_gh": ("classification300", M52_15, PROBIT, ("VI", dict()), 3, None),
    "probit300_m52_slep_ut": ("classification300", M52_15, PROBIT, ("SLEP", dict(intmethod="UT")), 3, None),
    "probit300_m52_ghks_damped": ("classification300", M52_15, PROBIT, ("GHKS", dict(damping=0.5)), 3, None),
    "logit300_m32_eks": ("classification300", M32_154, LOGIT, ("EKS", dict()), 3, None),
    "logit300_m32_eep1": ("classification300", M32_154, LOGIT, ("EEP", dict(power=1.0)), 3, None),
    "logit300_m32_vi_gh": ("classification300", M32_154, LOGIT, ("VI", dict()), 3, None),
    "logit300_m32_ghep1": ("classification300", M32_154, LOGIT, ("GHEP", dict(power=1)), 3, None),
    "c3_aircraft160_d59_poisson_ghep1": ("aircraft160", ("Sum", [AIR_M52, AIR_QP365, AIR_QP7]), POISSON, ("GHEP", dict(power=1)), 2, None),
    "c3_aircraft160_d31_poisson_vi_gh": ("aircraft160", ("Sum", [AIR_M52, AIR_QP365]), POISSON, ("VI", dict(intmethod="GH")), 2, None),
    "c3_aircraft160_d31_poisson_eks": ("aircraft160", ("Sum", [AIR_M52, AIR_QP365]), POISSON, ("EKS", dict()), 2, None),
    "k5_banana400_st52_probit_pep05": ("banana400", ST52, PROBIT, ("ExpectationPropagation", dict(power=0.5)), 2, 178.81),
    "c4_banana200_st52_probit_vi": ("banana200", ST52, PROBIT, ("VI", dict()), 2, None),
    # 25 inducing points -> state dimension 75: above what one SM's shared memory holds (global-scratch path)
    "c4_banana120_st52_m25_probit_ep": ("banana120", ST52_M25, PROBIT, ("EP", dict(power=0.8)), 2, None),
    # model zoo: the other closed-form priors, through the generic block-diagonal path
    "zoo_coal333_m12_poisson_ep05": ("coal333", M12_12, POISSON, ("EP", dict(power=0.5)), 3, None),
    "zoo_regression_m72_gauss_pep": ("regression", M72_15, ("Gaussian", dict(variance=0.5)), ("EP", dict(power=0.5)), 2, None),
    # (the reference's Cosine returns lists for L and Qc, so it cannot be a component of its Sum: used on its own)
    "zoo_regression_cosine_gauss_pep": ("regression", COS_05, ("Gaussian", dict(variance=0.5)), ("EP", dict(power=0.5)), 2, None),
    # (likewise Periodic: its 1-D Qc breaks the reference's Sum, priors.py:1376)
    "zoo_probit300_periodic_ep1": ("classification300", PER_3, PROBIT, ("EP", dict(power=1.0)), 3, None),
    "zoo_coal333_qpm12_poisson_vi_gh": ("coal333", QPM12, POISSON, ("VI", dict(intmethod="GH")), 3, None),
    "zoo_regression_sum_subbands_m12_gauss_ep1": ("regression", ("Sum", [SUB12, SUB32, M12_12]), ("Gaussian", dict(variance=0.5)),
                                                  ("EP", dict(power=1.0)), 2, None),
    "zoo_coal333_subband32_poisson_eks": ("coal333", SUB32, POISSON, ("EKS", dict(damping=0.8)), 3, None),
    # the 6 x 6 block kind (priors.py:605-721), on its own and next to narrower blocks
    "zoo_coal333_subband52_poisson_pep05": ("coal333", SUB52, POISSON, ("EP", dict(power=0.5)), 3, None),
    "zoo_regression_sum_subband52_m32_gauss_vi": ("regression", ("Sum", [SUB52, M32_154, SUB32]), ("Gaussian", dict(variance=0.5)),
                                                  ("VI", dict()), 2, None),
    "zoo_banana200_spatial52_probit_ep08": ("banana200", SP52, PROBIT, ("EP", dict(power=0.8)), 2, None),
    "zoo_banana200_spatial32_probit_vi": ("banana200", SP32, PROBIT, ("VI", dict()), 2, None),
    # multi-latent (func_dim = 2): Independent prior + HeteroscedasticNoise, the mcycle model of
    # notebooks/heteroscedastic_noise.py:51-58 (fold 0 of its 10-fold split)
    "het_mcycle_indep_m32_ep001_gh_damped": ("mcycle", ("Independent", [M32_31, M32_31]), HETERO,
                                             ("ExpectationPropagation", dict(power=0.01, intmethod="GH", damping=0.5)), 3, None),
    "het_mcycle_indep_m32_m52_ep09_ut_damped": ("mcycle", ("Independent", [M32_31, M52_205]), HETERO,
                                                ("ExpectationPropagation", dict(power=0.9, intmethod="UT", damping=0.1)), 3, None),
    "het_mcycle_indep_m32_vi_gh_damped": ("mcycle", ("Independent", [M32_31, M32_31]), HETERO,
                                          ("VI", dict(intmethod="GH", damping=0.5)), 3, None),
    "het_mcycle_indep_m32_m52_vi_ut": ("mcycle", ("Independent", [M32_31, M52_205]), HETERO,
                                       ("VI", dict(intmethod="UT")), 3, None),
    "het_mcycle_indep_m32_eep05_damped": ("mcycle", ("Independent", [M32_31, M32_31]), HETERO,
                                          ("EEP", dict(power=0.5, damping=0.5)), 3, None),
    "het_mcycle_indep_m32_m52_eks": ("mcycle", ("Independent", [M32_31, M52_205]), HETERO,
                                     ("EKS", dict()), 3, None),
    "het_mcycle_indep_m32_slep05_gh_damped": ("mcycle", ("Independent", [M32_31, M32_31]), HETERO,
                                              ("SLEP", dict(power=0.5, intmethod="GH", damping=0.5, num_cub_pts=10)), 3, None),
    "het_mcycle_indep_m32_m52_uks": ("mcycle", ("Independent", [M32_31, M52_205]), HETERO,
                                     ("UKS", dict()), 3, None),
    # (on its own the reference's Matern52FixedVar fails in softplus_list — its hyp is a scalar — so: inside a Sum)
    "zoo_coal333_sum_fixedvar_poisson_eks": ("coal333", ("Sum", [SUBFV, M52FV]),
                                             POISSON, ("EKS", dict()), 3, None),
}

KEEP_EVERY = {"aircraft160": 8, "banana400": 16, "banana200": 16, "banana120": 24}


def build(spec, module):
    """Instantiate `spec` from `module` (the reference's module, oracle.*, or the product's mirror)."""
    name, arg = spec
    if name in ("Sum", "Independent"):
        return getattr(module, name)([build(s, module) for s in arg])
    arg = dict(arg)
    if "z" in arg:
        import numpy as np
        arg["z"] = np.asarray(arg["z"], dtype=float)
    return getattr(module, name)(**arg)
