// Inert stand-in for <node.h> (and the libuv bits the reference glue names).
// TEST INFRASTRUCTURE ONLY -- see v8.h in this directory.
#ifndef PICHA_ORACLE_STUB_NODE_H
#define PICHA_ORACLE_STUB_NODE_H
#include "v8.h"

struct uv_loop_t;
struct uv_work_t { void* data; };
typedef void (*uv_work_cb)(uv_work_t*);
typedef void (*uv_after_work_cb)(uv_work_t*, int);
inline uv_loop_t* uv_default_loop() { return 0; }
inline int uv_queue_work(uv_loop_t*, uv_work_t*, uv_work_cb, uv_after_work_cb) { return 0; }

namespace node {}
#endif
