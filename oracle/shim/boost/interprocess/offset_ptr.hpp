// Test-infrastructure shim (oracle/Makefile, oracle/_ref only): lets UNMODIFIED reference sources compile without Boost.Interprocess (Boctree.h:37 uses offset_ptr only as a plain pointer outside the scan server).
#pragma once
namespace boost { namespace interprocess {
template <class T> class offset_ptr {
  T* p_;
 public:
  offset_ptr() : p_(nullptr) {}
  offset_ptr(T* p) : p_(p) {}
  offset_ptr& operator=(T* p) { p_ = p; return *this; }
  T* get() const { return p_; }
  T& operator*() const { return *p_; }
  T* operator->() const { return p_; }
  T& operator[](long i) const { return p_[i]; }
  operator T*() const { return p_; }
  explicit operator bool() const { return p_ != nullptr; }
  offset_ptr operator+(long i) const { return offset_ptr(p_ + i); }
  offset_ptr& operator+=(long i) { p_ += i; return *this; }
  offset_ptr& operator++() { ++p_; return *this; }
};
}}
