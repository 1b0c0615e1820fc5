// One family of instantiations of the N-body kernel (see nb_variants.hpp). Compiled several times with different
// -DHY_NB_LT / -DHY_NB_CTA.
#include "nb_variants.hpp"
#include "nb_kernel.cuh"

#if !defined(HY_NB_LT) || !defined(HY_NB_CTA)
#error "HY_NB_LT and HY_NB_CTA must be defined"
#endif

#define HY_NB_CAT_(a, b, c, d) a##b##c##d
#define HY_NB_CAT(a, b, c, d) HY_NB_CAT_(a, b, c, d)

namespace heyoka_b200::detail
{

namespace
{

#define HY_NB(TM, MAXT)                                                                                                \
    nb_variant                                                                                                         \
    {                                                                                                                  \
        HY_NB_LT, HY_NB_CTA != 0, TM, MAXT, dev::k_nb<HY_NB_LT, HY_NB_CTA != 0, TM, false, MAXT>,                      \
            dev::k_nb<HY_NB_LT, HY_NB_CTA != 0, TM, true, MAXT>                                                        \
    }

const nb_variant family[] = {
#if HY_NB_CTA != 0
    HY_NB(true, 512), HY_NB(false, 512)
#else
    HY_NB(true, 512), HY_NB(true, 384), HY_NB(true, 256), HY_NB(false, 512), HY_NB(false, 256)
#endif
};

} // namespace

nb_family HY_NB_CAT(nb_family_lt, HY_NB_LT, _cta, HY_NB_CTA)()
{
    return {family, sizeof(family) / sizeof(family[0])};
}

} // namespace heyoka_b200::detail
