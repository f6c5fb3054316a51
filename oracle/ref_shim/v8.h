// Inert stand-in for <v8.h>.  TEST INFRASTRUCTURE ONLY.
//
// The reference's hot path (resize.cc / colorconvert.cc) is plain C++ but sits in
// the same translation units as its NAN entry points.  These declarations exist
// only so that glue *parses*; nothing declared here is ever executed by the
// oracle (ref_unity.cc only calls picha::resizeImage / picha::doColorConvert).
#ifndef PICHA_ORACLE_STUB_V8_H
#define PICHA_ORACLE_STUB_V8_H
#include <stdint.h>
#include <stddef.h>
#include <algorithm>
#include <vector>

namespace v8 {

class Context; class Value; class Object; class String; class Function;
class Primitive; class Integer; class Array;

template <class T> class Local {
public:
	Local() : p_(0) {}
	template <class S> Local(const Local<S>&) : p_(0) {}
	T* operator->() const { return p_; }
	T* operator*() const { return p_; }
	bool IsEmpty() const { return true; }
	template <class S> static Local<T> Cast(const Local<S>&) { return Local<T>(); }
private:
	T* p_;
};

template <class T> class Maybe {
public:
	Maybe() : v_() {}
	T FromMaybe(const T& d) const { return d; }
private:
	T v_;
};

template <class T> class MaybeLocal {
public:
	MaybeLocal() {}
	template <class S> MaybeLocal(const Local<S>&) {}
	bool IsEmpty() const { return true; }
	Local<T> ToLocalChecked() const { return Local<T>(); }
	template <class S> Local<S> FromMaybe(const Local<S>& d) const { return d; }
	template <class S> bool ToLocal(Local<S>*) const { return false; }
};

class Context {
public:
	Local<Object> Global() { return Local<Object>(); }
};

class Value {
public:
	bool IsUndefined() const { return true; }
	bool IsObject() const { return false; }
	bool IsFunction() const { return false; }
	template <class S> bool StrictEquals(const Local<S>&) const { return false; }
	Maybe<double> NumberValue(Local<Context>) const { return Maybe<double>(); }
	Maybe<uint32_t> Uint32Value(Local<Context>) const { return Maybe<uint32_t>(); }
	MaybeLocal<Object> ToObject(Local<Context>) const { return MaybeLocal<Object>(); }
};

class Object : public Value {
public:
	template <class K> MaybeLocal<Value> Get(Local<Context>, const Local<K>&) { return MaybeLocal<Value>(); }
};
class Primitive : public Value {};
class String : public Primitive {};
class Integer : public Primitive {};
class Array : public Object {};
class Function : public Object {
public:
	template <class S> void SetName(const Local<S>&) {}
};

template <class T> class Persistent {
public:
	Persistent() {}
	void Reset() {}
	template <class S> void Reset(const Local<S>&) {}
private:
	Persistent(const Persistent&);
	void operator=(const Persistent&);
};

template <class T> class ReturnValue {
public:
	template <class S> void Set(const Local<S>&) {}
};

template <class T> class FunctionCallbackInfo {
public:
	int Length() const { return 0; }
	Local<Value> operator[](int) const { return Local<Value>(); }
	ReturnValue<T> GetReturnValue() const { return ReturnValue<T>(); }
};

}  // namespace v8
#endif
