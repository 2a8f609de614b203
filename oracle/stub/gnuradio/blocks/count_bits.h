/* TEST INFRASTRUCTURE ONLY -- gr::blocks::count_bits32 is a population count. */
#ifndef SNRX_STUB_GR_COUNT_BITS_H
#define SNRX_STUB_GR_COUNT_BITS_H
namespace gr { namespace blocks {
inline unsigned int count_bits32(unsigned int x) { return (unsigned int)__builtin_popcount(x); }
} }
#endif
