// stand-in for <boost/circular_buffer.hpp> (lib/decoder_impl.h:32,104): fixed capacity, push_back
// overwrites the oldest element when full, operator[] indexes from the oldest element.
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
template <class T>
class circular_buffer {
public:
    explicit circular_buffer(size_t capacity) : d_buf(capacity), d_first(0), d_size(0) {}
    void push_back(const T &v) {
        const size_t cap = d_buf.size();
        if (cap == 0) return;
        if (d_size < cap) {
            d_buf[(d_first + d_size) % cap] = v;
            d_size++;
        } else {
            d_buf[d_first] = v;
            d_first = (d_first + 1) % cap;
        }
    }
    size_t size() const { return d_size; }
    size_t capacity() const { return d_buf.size(); }
    T &operator[](size_t i) { return d_buf[(d_first + i) % d_buf.size()]; }
    const T &operator[](size_t i) const { return d_buf[(d_first + i) % d_buf.size()]; }
    T &front() { return (*this)[0]; }
    T &back() { return (*this)[d_size - 1]; }
    void clear() { d_first = 0; d_size = 0; }
private:
    std::vector<T> d_buf;
    size_t d_first, d_size;
};
}  // namespace boost
