// heyoka_b200 — named arguments (kw::tol = 1e-12, kw::masses = {1., 0.}, ...).
//
// The reference parses named arguments with the igor library (include/heyoka/detail/igor.hpp,
// include/heyoka/kw.hpp:30-130). Only the *surface* is kept here: the same names, usable with the
// same `kw::name = value` syntax, resolved with a few lines of template code.
#ifndef HEYOKA_B200_KW_HPP
#define HEYOKA_B200_KW_HPP

#include <initializer_list>
#include <type_traits>
#include <utility>
#include <vector>

namespace heyoka_b200
{

namespace kw
{

namespace detail
{

template <typename Tag, typename T>
struct tagged_arg {
    using tag = Tag;
    T value;
};

template <typename Tag>
struct named {
    template <typename T>
    constexpr auto operator=(T &&v) const
    {
        return tagged_arg<Tag, std::decay_t<T>>{std::forward<T>(v)};
    }
    template <typename T>
    auto operator=(std::initializer_list<T> il) const
    {
        return tagged_arg<Tag, std::vector<T>>{std::vector<T>(il)};
    }
};

template <typename T>
struct is_tagged : std::false_type {
};
template <typename Tag, typename T>
struct is_tagged<tagged_arg<Tag, T>> : std::true_type {
};

template <typename Tag, typename Arg>
constexpr bool matches()
{
    if constexpr (is_tagged<std::decay_t<Arg>>::value) {
        return std::is_same_v<typename std::decay_t<Arg>::tag, Tag>;
    } else {
        return false;
    }
}

} // namespace detail

// True if one of the arguments carries Tag.
template <typename Tag, typename... Args>
constexpr bool has(const detail::named<Tag> &, const Args &...)
{
    return (detail::matches<Tag, Args>() || ... || false);
}

template <typename Tag, typename... Args>
constexpr bool has_tag()
{
    return (detail::matches<Tag, Args>() || ... || false);
}

// Apply f to the value tagged with Tag, if present.
template <typename Tag, typename F, typename... Args>
void visit(const detail::named<Tag> &, F &&f, const Args &...args)
{
    const auto one = [&f](const auto &a) {
        if constexpr (detail::matches<Tag, decltype(a)>()) {
            f(a.value);
        }
    };
    (void)one; // (an empty pack leaves it unused)
    (one(args), ...);
}

// True if every argument is a named argument whose tag is in the Allowed list.
template <typename... Allowed>
struct allowed_tags {
    template <typename Arg>
    static constexpr bool ok()
    {
        if constexpr (detail::is_tagged<std::decay_t<Arg>>::value) {
            return (std::is_same_v<typename std::decay_t<Arg>::tag, Allowed> || ...);
        } else {
            return false;
        }
    }
    template <typename... Args>
    static constexpr bool all()
    {
        return (ok<Args>() && ... && true);
    }
};

#define HEYOKA_B200_KWARG(name)                                                                                        \
    struct name##_tag {                                                                                                \
    };                                                                                                                 \
    inline constexpr detail::named<name##_tag> name {}

// Integrator construction (include/heyoka/kw.hpp, include/heyoka/taylor.hpp:178-181,814-821).
HEYOKA_B200_KWARG(tol);
HEYOKA_B200_KWARG(high_accuracy);
HEYOKA_B200_KWARG(compact_mode);
HEYOKA_B200_KWARG(pars);
HEYOKA_B200_KWARG(time);
HEYOKA_B200_KWARG(parallel_mode);
HEYOKA_B200_KWARG(parjit);
HEYOKA_B200_KWARG(t_events);
HEYOKA_B200_KWARG(nt_events);
// llvm_state options: accepted and ignored (there is no JIT).
HEYOKA_B200_KWARG(opt_level);
HEYOKA_B200_KWARG(fast_math);
HEYOKA_B200_KWARG(force_avx512);
HEYOKA_B200_KWARG(slp_vectorize);
HEYOKA_B200_KWARG(mname);
HEYOKA_B200_KWARG(code_model);
// Propagation (include/heyoka/taylor.hpp:271-272,719-729).
HEYOKA_B200_KWARG(max_steps);
HEYOKA_B200_KWARG(max_delta_t);
HEYOKA_B200_KWARG(callback);
HEYOKA_B200_KWARG(write_tc);
HEYOKA_B200_KWARG(c_output);
// Events (include/heyoka/events.hpp:49, include/heyoka/kw.hpp).
HEYOKA_B200_KWARG(cooldown);
HEYOKA_B200_KWARG(direction);
// Models.
HEYOKA_B200_KWARG(Gconst);
HEYOKA_B200_KWARG(masses);
HEYOKA_B200_KWARG(gconst);
HEYOKA_B200_KWARG(length);
HEYOKA_B200_KWARG(inputs);
HEYOKA_B200_KWARG(nn_hidden);
HEYOKA_B200_KWARG(n_out);
HEYOKA_B200_KWARG(activations);
HEYOKA_B200_KWARG(nn_wb);
// Device selection (extension: which GPU owns the batch; default: the current device).
HEYOKA_B200_KWARG(device);
// Sharding (extension): the GPUs the lanes of the batch are split over (contiguous blocks, hy_batch_create_multi()).
HEYOKA_B200_KWARG(devices);

#undef HEYOKA_B200_KWARG

} // namespace kw

} // namespace heyoka_b200

#endif
