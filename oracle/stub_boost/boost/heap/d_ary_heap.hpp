// TEST INFRASTRUCTURE ONLY -- stand-in for boost::heap::d_ary_heap, just enough of its interface to compile the reference's
// thirdparty/jps3d/src/jps_planner/graph_search.cpp (which needs nothing else beyond the standard library) into
// oracle/_ref/.  Boost is not installed in this image.  Semantics kept: a max-heap with respect to Compare (top() is an
// element that no other element compares greater than), mutable handles, increase() after the priority of an element went
// up.  NOT kept: boost's internal tie-breaking order among equal-priority elements (unspecified by boost), so paths found
// through this stub may differ from a boost build among equal-cost alternatives; costs cannot.
#pragma once
#include <cstddef>
#include <cstdio>   // the reference relies on boost pulling this in (printf in graph_search.cpp)
#include <memory>
#include <utility>
#include <vector>

namespace boost
{
namespace heap
{
template <bool B>
struct mutable_
{
};
template <int N>
struct arity
{
};
template <class C>
struct compare
{
  typedef C type;
};

template <class T, class A0, class A1, class Cmp>
class d_ary_heap
{
  struct Node
  {
    T value;
    size_t pos;
  };

public:
  class handle_type
  {
  public:
    handle_type() {}
    std::shared_ptr<Node> n;
  };
  typedef typename Cmp::type Compare;

  handle_type push(const T& v)
  {
    handle_type h;
    h.n = std::make_shared<Node>();
    h.n->value = v;
    h.n->pos = a_.size();
    a_.push_back(h.n);
    up(a_.size() - 1);
    return h;
  }
  const T& top() const { return a_[0]->value; }
  void pop()
  {
    a_[0] = a_.back();
    a_[0]->pos = 0;
    a_.pop_back();
    if (!a_.empty()) down(0);
  }
  void increase(const handle_type& h) { up(h.n->pos); }
  void update(const handle_type& h) { up(h.n->pos); down(h.n->pos); }
  bool empty() const { return a_.empty(); }
  size_t size() const { return a_.size(); }
  void clear() { a_.clear(); }

  // ordered iteration is only used by getOpenSet() (visualisation); plain storage order suffices
  class const_iterator
  {
  public:
    typename std::vector<std::shared_ptr<Node>>::const_iterator it;
    const T& operator*() const { return (*it)->value; }
    const_iterator& operator++() { ++it; return *this; }
    bool operator!=(const const_iterator& o) const { return it != o.it; }
  };
  typedef const_iterator ordered_iterator;
  const_iterator begin() const { return const_iterator{ a_.begin() }; }
  const_iterator end() const { return const_iterator{ a_.end() }; }
  const_iterator ordered_begin() const { return begin(); }
  const_iterator ordered_end() const { return end(); }

private:
  // cmp_(a, b) == true  <=>  a has LOWER priority than b
  void up(size_t i)
  {
    while (i > 0)
    {
      const size_t p = (i - 1) / 2;
      if (!cmp_(a_[p]->value, a_[i]->value)) break;
      std::swap(a_[p], a_[i]);
      a_[p]->pos = p; a_[i]->pos = i;
      i = p;
    }
  }
  void down(size_t i)
  {
    const size_t n = a_.size();
    for (;;)
    {
      size_t best = i;
      const size_t l = 2 * i + 1, r = l + 1;
      if (l < n && cmp_(a_[best]->value, a_[l]->value)) best = l;
      if (r < n && cmp_(a_[best]->value, a_[r]->value)) best = r;
      if (best == i) break;
      std::swap(a_[best], a_[i]);
      a_[best]->pos = best; a_[i]->pos = i;
      i = best;
    }
  }
  std::vector<std::shared_ptr<Node>> a_;
  Compare cmp_;
};
}  // namespace heap
}  // namespace boost
