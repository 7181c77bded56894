// havoqgt/visitor_queue.hpp — the visitor queue for ONE rank: the control flow of the reference's
// init_visitor_traversal_new() and queue_visitor() (include/havoqgt/visitor_queue.hpp:221-251, 395-411) with the mailbox,
// the delegate broadcast and the termination detection taken out (one rank owns every vertex, nothing is a delegate):
//   * queue_visitor(v): v.pre_visit() decides whether v enters the local queue (:399-402)
//   * traversal: one init_visit per vertex in vertex order, and after EACH of them the local queue is drained
//     (pop, visit) before the next vertex starts (:225-242)
// The local queue is the Queue template the caller names — the reference's own visitor_priority_queue or tppm_queue.
#pragma once
#include <havoqgt/environment.hpp>
#include <havoqgt/mpi.hpp>

namespace havoqgt {
namespace mpi {

class oned_blocked_partitioned_t {};
class el_partitioned_t {};

template <typename TVisitor, template <typename T> class Queue, typename TGraph, typename AlgData>
class visitor_queue {
 public:
  typedef TVisitor visitor_type;
  typedef typename TGraph::vertex_locator vertex_locator;

  visitor_queue(TGraph* g, AlgData& alg_data) : m_ptr_graph(g), m_alg_data(alg_data) {}

  void init_visitor_traversal_new() {
    for (auto vitr = m_ptr_graph->vertices_begin(); vitr != m_ptr_graph->vertices_end(); ++vitr) {
      visitor_type v(*vitr);
      v.init_visit(*m_ptr_graph, this, m_alg_data);
      drain();
    }
    drain();
  }

  void queue_visitor(const visitor_type& v) {
    if (v.pre_visit(m_alg_data)) m_local_queue.push(v);
  }

 private:
  void drain() {
    while (!m_local_queue.empty()) {
      visitor_type v = m_local_queue.top();
      m_local_queue.pop();
      v.visit(*m_ptr_graph, this, m_alg_data);
    }
  }

  Queue<visitor_type> m_local_queue;
  TGraph* m_ptr_graph;
  AlgData& m_alg_data;
};

template <typename TVisitor, template <typename T> class Queue, typename TGraph, typename AlgData>
visitor_queue<TVisitor, Queue, TGraph, AlgData> create_visitor_queue(TGraph* g, AlgData& alg_data) {
  return visitor_queue<TVisitor, Queue, TGraph, AlgData>(g, alg_data);
}

}  // namespace mpi
}  // namespace havoqgt
