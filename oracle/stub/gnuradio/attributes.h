/* TEST INFRASTRUCTURE ONLY -- stand-in for <gnuradio/attributes.h> (oracle build). */
#ifndef SNRX_STUB_GR_ATTRIBUTES_H
#define SNRX_STUB_GR_ATTRIBUTES_H
#define __GR_ATTR_EXPORT
#define __GR_ATTR_IMPORT
#endif
