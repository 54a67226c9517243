// ORACLE / TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN event selector: lotto::RejectionFreeEventSelector
// (submodules/kmc-lotto/include/lotto/rejection_free.hpp:50-198, header-only), compiled
// unmodified from where it lies under /root/reference into oracle/_ref/libkmc_lotto.so
// (oracle/Makefile, target `lotto`).  This file only supplies what the reference's
// CompleteKineticEventData supplies around it (monte_calculator/kinetic_events.hh:73-133):
// a rate calculator (here: a callback into the test, which evaluates the event state with
// the reference's generated Clexulator kernels), the complete event id list in the order
// of make_complete_event_id_list (events/CompleteEventList.cc:76-91), the impact table,
// and a seeded std::mt19937_64 (default_engine_type).
//
// Imported only by tests/ and tools/; never by the product.
#include <cstdint>
#include <map>
#include <memory>
#include <random>
#include <vector>

#include "lotto/rejection_free.hpp"

namespace {

typedef double (*rate_cb)(void *ctx, long event_id);

struct RateCalc {
  rate_cb cb;
  void *ctx;
  double calculate_rate(long const &event_id) const { return cb(ctx, event_id); }
};

typedef lotto::RejectionFreeEventSelector<long, RateCalc, std::mt19937_64> Selector;

struct Handle {
  std::shared_ptr<RateCalc> calc;
  std::shared_ptr<Selector> selector;
};

}  // namespace

extern "C" {

// imp_beg[n_events + 1], imp[...]: impacted event ids per event
void *lotto_create(long n_events, rate_cb cb, void *ctx, long const *imp_beg, long const *imp,
                   unsigned long long seed) {
  auto h = new Handle;
  h->calc = std::make_shared<RateCalc>(RateCalc{cb, ctx});
  std::vector<long> ids(n_events);
  std::map<long, std::vector<long>> table;
  for (long e = 0; e < n_events; ++e) {
    ids[e] = e;
    table[e] = std::vector<long>(imp + imp_beg[e], imp + imp_beg[e + 1]);
  }
  auto engine = std::make_shared<std::mt19937_64>();
  engine->seed(seed);
  auto rng = std::make_shared<lotto::RandomGeneratorT<std::mt19937_64>>(engine);
  h->selector = std::make_shared<Selector>(h->calc, ids, table, rng);
  return h;
}

void lotto_destroy(void *p) { delete static_cast<Handle *>(p); }

// one select_event(): returns the event id; *dt the time step, *total the total rate used
long lotto_select(void *p, double *dt, double *total) {
  Handle *h = static_cast<Handle *>(p);
  std::pair<long, double> sel = h->selector->select_event();
  if (dt) *dt = sel.second;
  if (total) *total = h->selector->total_rate();
  return sel.first;
}

// n_steps of the loop of kinetic_monte_carlo_v2 (methods/kinetic_monte_carlo.hh:396-507):
// select, advance the time, apply (callback).  Returns the simulated time; event_log[n_steps]
// (may be null) receives the selected event ids.
typedef void (*apply_cb)(void *ctx, long event_id);
double lotto_run(void *p, long n_steps, apply_cb apply, void *ctx, long *event_log) {
  Handle *h = static_cast<Handle *>(p);
  double time = 0.0;
  for (long s = 0; s < n_steps; ++s) {
    std::pair<long, double> sel = h->selector->select_event();
    time += sel.second;
    apply(ctx, sel.first);
    if (event_log) event_log[s] = sel.first;
  }
  return time;
}

double lotto_get_rate(void *p, long event_id) {
  return static_cast<Handle *>(p)->selector->get_rate(event_id);
}
}
