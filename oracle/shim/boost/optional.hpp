#pragma once
#include <cassert>
#include <utility>
namespace boost {
struct none_t {}; static const none_t none = none_t();
template<typename T> class optional {
  bool has_; T val_;
public:
  optional() : has_(false), val_() {}
  optional(none_t) : has_(false), val_() {}
  optional(const T& v) : has_(true), val_(v) {}
  optional& operator=(none_t) { has_ = false; val_ = T(); return *this; }
  optional& operator=(const T& v) { has_ = true; val_ = v; return *this; }
  explicit operator bool() const { return has_; }
  bool operator!() const { return !has_; }
  T& operator*() { return val_; } const T& operator*() const { return val_; }
  T* operator->() { return &val_; } const T* operator->() const { return &val_; }
  T& get() { return val_; } const T& get() const { return val_; }
};
}
