/* TEST INFRASTRUCTURE ONLY -- the smallest gr::block / pmt / boost surface that lets the
 * reference scapy-radio/gnuradio/gr-zigbee/lib/packet_sink_scapy_impl.cc compile unmodified.
 * Nothing of GNU Radio's scheduler is modelled: general_work() is called directly by
 * oracle/zb_ref_harness.cc and message_port_pub() records the published blob. */
#ifndef SNRX_STUB_GR_BLOCK_H
#define SNRX_STUB_GR_BLOCK_H
#include <cassert>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

typedef std::vector<int> gr_vector_int;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }

namespace pmt {
struct pmt_base {
    std::string sym;
    std::vector<uint8_t> blob;
    std::shared_ptr<pmt_base> car, cdr;
};
typedef std::shared_ptr<pmt_base> pmt_t;
inline pmt_t mp(const char* s) { auto p = std::make_shared<pmt_base>(); p->sym = s; return p; }
inline pmt_t make_dict() { return std::make_shared<pmt_base>(); }
inline pmt_t make_blob(const void* d, size_t n) {
    auto p = std::make_shared<pmt_base>();
    p->blob.assign((const uint8_t*)d, (const uint8_t*)d + n);
    return p;
}
inline pmt_t cons(pmt_t a, pmt_t b) { auto p = std::make_shared<pmt_base>(); p->car = a; p->cdr = b; return p; }
}  // namespace pmt

namespace gr {
class io_signature {
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int, int, int) { return std::make_shared<io_signature>(); }
};
class block {
public:
    block() {}
    block(const std::string&, io_signature::sptr, io_signature::sptr) {}
    virtual ~block() {}
    std::vector<std::vector<uint8_t>> published;       /* blobs in publication order */
    void message_port_register_out(pmt::pmt_t) {}
    void message_port_pub(pmt::pmt_t, pmt::pmt_t msg) { published.push_back(msg->cdr->blob); }
    void consume(int, int) {}
    virtual int general_work(int, gr_vector_int&, gr_vector_const_void_star&, gr_vector_void_star&) { return 0; }
};
}  // namespace gr

namespace gnuradio {
template <class T> std::shared_ptr<T> get_initial_sptr(T* p) { return std::shared_ptr<T>(p); }
}
#endif
