// Test-infrastructure only. Minimal stand-in for the single oneTBB entry point the reference
// uses (tbb::detail::d1::parallel_for(int,int,lambda), reference source/functions.cpp:100,119,367,493),
// so the UNMODIFIED reference sources compile where they lie (oneTBB is not installed here).
#pragma once
namespace tbb { namespace detail { namespace d1 {
template <typename Index, typename Body>
void parallel_for(Index first, Index last, const Body& body) {
#pragma omp parallel for schedule(dynamic, 8)
  for (Index i = first; i < last; ++i) body(i);
}
}}}  // namespace tbb::detail::d1
