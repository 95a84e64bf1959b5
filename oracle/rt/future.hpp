// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into libmsim_b200.so; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Poll-based futures for the CPU restatement of madsim's executor.
//
// Rust futures are *pulled*: a wake re-polls the task's root future, which re-polls the
// chain of sub-futures down to the leaf (this is observable in madsim: `Sleep::poll`
// registers a NEW timer on every pending poll, time/sleep.rs:47-54).  C++20 coroutines are
// *pushed* (resume at the suspension point), so `Async<T>` below re-creates pull semantics:
// a suspended coroutine remembers the leaf it awaits, and `Async::poll` first re-polls that
// leaf and only resumes the body once the leaf is Ready.
#pragma once
#include <coroutine>
#include <cstdint>
#include <exception>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

namespace msim_oracle {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i64 = int64_t;

struct Unit {};

/// A Rust panic.  Unwinds out of `poll` to the executor (task/mod.rs:285-317).
struct Panic : std::runtime_error {
  u32 code;
  explicit Panic(const std::string& m, u32 code_ = 0) : std::runtime_error(m), code(code_) {}
};
[[noreturn]] inline void panic(const std::string& m, u32 code = 0) { throw Panic(m, code); }

template <class T>
struct Poll {
  std::optional<T> v;
  using value_type = T;
  bool ready() const { return v.has_value(); }
  static Poll pending() { return Poll{}; }
  static Poll done(T x) {
    Poll p;
    p.v.emplace(std::move(x));
    return p;
  }
};

struct RawTask;  // task.hpp
/// std::task::Waker: a counted reference to the task (async-task's RawWaker).
struct Waker {
  std::shared_ptr<RawTask> task;  // null = noop waker
  void wake() const;              // defined in task.hpp
  bool will_wake(const Waker& o) const { return task == o.task; }
};
struct Context {
  Waker waker;
};

// ---------------------------------------------------------------------------------------
// Async<T>: a coroutine that IS a pollable future.
template <class F, bool Owned>
struct PollAwaiter;

struct PromiseBase {
  Context* cx = nullptr;                         // valid only during a poll
  void* leaf = nullptr;                          // awaiter of the pending sub-future
  bool (*leaf_poll)(void*, Context&) = nullptr;  // re-poll it; true when Ready
  std::exception_ptr ex;
  template <class F>
  auto await_transform(F&& f);
};

template <class F>
using poll_value_t = typename decltype(std::declval<F&>().poll(std::declval<Context&>()))::value_type;

template <class F, bool Owned>
struct PollAwaiter {
  using R = poll_value_t<F>;
  std::conditional_t<Owned, F, F*> fut;
  PromiseBase* p;
  std::optional<R> result;
  F& f() {
    if constexpr (Owned)
      return fut;
    else
      return *fut;
  }
  bool try_poll(Context& cx) {
    auto r = f().poll(cx);
    if (r.ready()) {
      result = std::move(r.v);
      return true;
    }
    return false;
  }
  bool await_ready() { return try_poll(*p->cx); }
  void await_suspend(std::coroutine_handle<>) {
    p->leaf = this;
    p->leaf_poll = [](void* self, Context& cx) { return static_cast<PollAwaiter*>(self)->try_poll(cx); };
  }
  R await_resume() { return std::move(*result); }
};

template <class F>
auto PromiseBase::await_transform(F&& f) {
  using D = std::decay_t<F>;
  if constexpr (std::is_lvalue_reference_v<F>)
    return PollAwaiter<D, false>{&f, this, std::nullopt};
  else
    return PollAwaiter<D, true>{std::move(f), this, std::nullopt};
}

template <class T>
class [[nodiscard]] Async {
 public:
  struct promise_type : PromiseBase {
    std::optional<T> value;
    Async get_return_object() { return Async(std::coroutine_handle<promise_type>::from_promise(*this)); }
    std::suspend_always initial_suspend() noexcept { return {}; }
    std::suspend_always final_suspend() noexcept { return {}; }
    void return_value(T v) { value.emplace(std::move(v)); }
    void unhandled_exception() { this->ex = std::current_exception(); }
  };
  using handle_t = std::coroutine_handle<promise_type>;
  using value_type = T;

  Async() = default;
  explicit Async(handle_t h) : h_(h) {}
  Async(Async&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  Async& operator=(Async&& o) noexcept {
    if (this != &o) {
      reset();
      h_ = std::exchange(o.h_, nullptr);
    }
    return *this;
  }
  Async(const Async&) = delete;
  Async& operator=(const Async&) = delete;
  ~Async() { reset(); }
  /// Dropping a future: destroys the frame at its suspension point (runs the destructors
  /// of live locals, e.g. a oneshot::Receiver or a BindGuard).
  void reset() {
    if (h_) {
      h_.destroy();
      h_ = nullptr;
    }
  }
  bool valid() const { return bool(h_); }

  Poll<T> poll(Context& cx) {
    auto& p = h_.promise();
    p.cx = &cx;
    if (p.leaf) {
      if (!p.leaf_poll(p.leaf, cx)) return Poll<T>::pending();
      p.leaf = nullptr;
    }
    h_.resume();
    p.cx = nullptr;
    if (h_.done()) {
      if (p.ex) {
        auto e = p.ex;
        p.ex = nullptr;
        std::rethrow_exception(e);
      }
      return Poll<T>::done(std::move(*p.value));
    }
    return Poll<T>::pending();
  }

 private:
  handle_t h_{};
};

/// async_stream::stream! { .. }: a generator that may both `co_await` futures and `co_yield`
/// items; `poll_next` is futures::Stream::poll_next (Ready(None) once the body returns).
template <class T>
class [[nodiscard]] Stream {
 public:
  struct promise_type : PromiseBase {
    std::optional<T> yielded;
    Stream get_return_object() { return Stream(std::coroutine_handle<promise_type>::from_promise(*this)); }
    std::suspend_always initial_suspend() noexcept { return {}; }
    std::suspend_always final_suspend() noexcept { return {}; }
    std::suspend_always yield_value(T v) {
      yielded.emplace(std::move(v));
      return {};
    }
    void return_void() {}
    void unhandled_exception() { this->ex = std::current_exception(); }
  };
  using handle_t = std::coroutine_handle<promise_type>;
  Stream() = default;
  explicit Stream(handle_t h) : h_(h) {}
  Stream(Stream&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  Stream& operator=(Stream&& o) noexcept {
    if (this != &o) {
      reset();
      h_ = std::exchange(o.h_, nullptr);
    }
    return *this;
  }
  Stream(const Stream&) = delete;
  ~Stream() { reset(); }
  void reset() {
    if (h_) {
      h_.destroy();
      h_ = nullptr;
    }
  }
  Poll<std::optional<T>> poll_next(Context& cx) {
    using R = Poll<std::optional<T>>;
    if (!h_ || h_.done()) return R::done(std::nullopt);
    auto& p = h_.promise();
    p.cx = &cx;
    if (p.leaf) {
      if (!p.leaf_poll(p.leaf, cx)) return R::pending();
      p.leaf = nullptr;
    }
    h_.resume();
    p.cx = nullptr;
    if (p.yielded) {
      std::optional<T> v = std::move(p.yielded);
      p.yielded.reset();
      return R::done(std::move(v));
    }
    if (h_.done()) {
      if (p.ex) {
        auto e = p.ex;
        p.ex = nullptr;
        std::rethrow_exception(e);
      }
      return R::done(std::nullopt);
    }
    return R::pending();
  }

 private:
  handle_t h_{};
};
/// StreamExt::next()
template <class T>
struct StreamNext {
  Stream<T>* s;
  Poll<std::optional<T>> poll(Context& cx) { return s->poll_next(cx); }
};

/// std::future::pending()
template <class T = Unit>
struct PendingForever {
  Poll<T> poll(Context&) { return Poll<T>::pending(); }
};

/// std::future::ready(v)
template <class T>
struct ReadyNow {
  T v;
  Poll<T> poll(Context&) { return Poll<T>::done(std::move(v)); }
};

/// tokio::task::yield_now(): Pending once (waking itself), then Ready.
struct YieldNow {
  bool yielded = false;
  Poll<Unit> poll(Context& cx) {
    if (yielded) return Poll<Unit>::done({});
    yielded = true;
    cx.waker.wake();
    return Poll<Unit>::pending();
  }
};

/// Result<T, E> with E an integer error kind (io::ErrorKind subset).
enum class Err : int {
  None = 0,
  AddrNotAvailable,
  AddrInUse,
  InvalidInput,
  ConnectionRefused,
  ConnectionReset,
  BrokenPipe,
  NotConnected,
  TimedOut,
  Elapsed,
  Cancelled,
};
template <class T>
struct Result {
  std::optional<T> ok;
  Err err = Err::None;
  static Result Ok(T v) {
    Result r;
    r.ok.emplace(std::move(v));
    return r;
  }
  static Result Error(Err e) {
    Result r;
    r.err = e;
    return r;
  }
  bool is_ok() const { return ok.has_value(); }
  bool is_err() const { return !ok.has_value(); }
  T& unwrap() {
    if (!ok) panic("called `Result::unwrap()` on an `Err` value: " + std::to_string(int(err)));
    return *ok;
  }
};

}  // namespace msim_oracle
