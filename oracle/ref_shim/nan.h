// Inert stand-in for <nan.h>.  TEST INFRASTRUCTURE ONLY -- see v8.h here.
#ifndef PICHA_ORACLE_STUB_NAN_H
#define PICHA_ORACLE_STUB_NAN_H
#include "node.h"

namespace Nan {

template <class T> class Persistent : public v8::Persistent<T> {};

typedef const v8::FunctionCallbackInfo<v8::Value>& NAN_METHOD_ARGS_TYPE;

class HandleScope { public: HandleScope() {} };
class TryCatch { public: TryCatch() {} bool HasCaught() const { return false; } };
inline void FatalException(const TryCatch&) {}

class AsyncResource {
public:
	explicit AsyncResource(const char*) {}
	template <class R, class F, class A> void runInAsyncScope(const R&, const F&, int, A*) {}
};

inline v8::Local<v8::Primitive> Undefined() { return v8::Local<v8::Primitive>(); }
inline v8::Local<v8::Context> GetCurrentContext() { return v8::Local<v8::Context>(); }
inline void ThrowError(const char*) {}
inline v8::Local<v8::Value> Error(const char*) { return v8::Local<v8::Value>(); }

template <class T> v8::Local<T> New(const Persistent<T>&) { return v8::Local<T>(); }
template <class T> v8::Local<T> New(const v8::Persistent<T>&) { return v8::Local<T>(); }
template <class T> v8::Local<T> New(const v8::Local<T>& l) { return l; }
inline v8::MaybeLocal<v8::String> New(const char*) { return v8::MaybeLocal<v8::String>(); }

template <class K> v8::MaybeLocal<v8::Value> Get(const v8::Local<v8::Object>&, const v8::Local<K>&) {
	return v8::MaybeLocal<v8::Value>();
}

}  // namespace Nan

#define NAN_METHOD(name) void name(Nan::NAN_METHOD_ARGS_TYPE info)
#endif
