// TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for btllib's <btllib/order_queue.hpp>, so the
// reference sources compile unmodified into oracle/_ref/. Surface used by the reference:
//   goldrush_path/read_hashing.cpp:17-19,57-71 (Block, write), goldrush_path.cpp:1210-1237 (read).
// Contract restated: blocks are delivered to read() strictly in increasing Block::num order,
// whatever order the producers write() them in; at most queue_size blocks are buffered.
#ifndef GRB_SHIM_BTLLIB_ORDER_QUEUE_HPP
#define GRB_SHIM_BTLLIB_ORDER_QUEUE_HPP

#include <condition_variable>
#include <cstddef>
#include <mutex>
#include <utility>
#include <vector>

namespace btllib {

template<typename T>
class OrderQueueMPMC
{
public:
  struct Block
  {
    explicit Block(size_t block_size)
      : data(block_size)
    {}
    Block(const Block&) = default;
    Block(Block&&) = default;
    Block& operator=(const Block&) = default;
    Block& operator=(Block&&) = default;

    std::vector<T> data;
    size_t count = 0;
    size_t num = 0;
  };

  OrderQueueMPMC(size_t queue_size, size_t block_size)
    : m_queue_size(queue_size)
    , m_block_size(block_size)
    , m_occupied(queue_size, false)
  {
    m_slots.reserve(queue_size);
    for (size_t i = 0; i < queue_size; ++i) {
      m_slots.emplace_back(block_size);
    }
  }

  // Moves `block` in; leaves the caller's block with storage for block_size items again.
  void write(Block& block)
  {
    const size_t num = block.num;
    std::unique_lock<std::mutex> lock(m_mutex);
    m_changed.wait(lock, [&] {
      return num < m_next_read + m_queue_size && !m_occupied[num % m_queue_size];
    });
    Block fresh(m_block_size);
    std::swap(m_slots[num % m_queue_size], block);
    block = std::move(fresh);
    m_occupied[num % m_queue_size] = true;
    m_changed.notify_all();
  }

  void read(Block& block)
  {
    std::unique_lock<std::mutex> lock(m_mutex);
    const size_t slot = m_next_read % m_queue_size;
    m_changed.wait(lock, [&] { return bool(m_occupied[slot]); });
    std::swap(block, m_slots[slot]);
    m_occupied[slot] = false;
    ++m_next_read;
    m_changed.notify_all();
  }

private:
  size_t m_queue_size;
  size_t m_block_size;
  std::vector<Block> m_slots;
  std::vector<char> m_occupied;
  size_t m_next_read = 0;
  std::mutex m_mutex;
  std::condition_variable m_changed;
};

} // namespace btllib

#endif
