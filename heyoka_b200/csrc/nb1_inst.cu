// The instantiations of the one-thread-per-lane N-body kernel (nb1_kernel.cuh), see nb_variants.hpp.
#include "nb_variants.hpp"
#include "nb1_kernel.cuh"

namespace heyoka_b200::detail
{

namespace
{

#define HY_NB1(TM, MAXT)                                                                                               \
    nb_variant                                                                                                         \
    {                                                                                                                  \
        32, false, TM, MAXT, dev::k_nb1<TM, false, MAXT>, dev::k_nb1<TM, true, MAXT>, true                             \
    }

const nb_variant family[] = {HY_NB1(true, 512), HY_NB1(true, 384), HY_NB1(true, 256), HY_NB1(false, 512), HY_NB1(false, 256)};

} // namespace

nb_family nb_family_lane()
{
    return {family, sizeof(family) / sizeof(family[0])};
}

} // namespace heyoka_b200::detail
