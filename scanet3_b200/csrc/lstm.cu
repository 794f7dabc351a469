#include <stdexcept>
#include "engine.h"
#include "tc.h"
namespace sb {
void lstm_plan(sb_engine*) {}
void lstm_forward(sb_engine*, LayerPlan&, int) { throw std::runtime_error("LSTM device path not built yet"); }
void lstm_backward(sb_engine*, LayerPlan&, int) { throw std::runtime_error("LSTM device path not built yet"); }
void lstm_free(sb_engine*) {}
}
