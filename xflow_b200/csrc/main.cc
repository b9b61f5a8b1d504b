// xflow_lr — the reference's binary (src/model/main.cc:15-48) on the drop-in: same argv
//   xflow_lr <train_prefix> <test_prefix> <model: 0 = LR, 1 = FM> <epochs>
// One process per GPU; there are no scheduler / server processes to start (the table lives on the GPU
// of this process).  Optimizer: env XFLOW_OPTIMIZER=ftrl|sgd (default ftrl, like server.h:24,28).
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "../../include/xflow/xflow.h"

int main(int argc, char* argv[]) {
  if (argc != 5) {
    std::cout << "usage: xflow_lr train_prefix test_prefix model_index epochs\n";
    std::cout << "LR model example: xflow_lr data/small_train data/small_test 0 100\n";
    std::cout << "FM model example: xflow_lr data/small_train data/small_test 1 100\n";
    return 2;
  }
  try {
    xflow::Server::Get();  // the reference builds the Server before ps::Start (main.cc:22-25)
    int epochs = std::atoi(argv[4]);
    if (*(argv[3]) == '0') {
      std::cout << "start LR " << std::endl;
      xflow::LRWorker lr_worker(argv[1], argv[2]);
      lr_worker.epochs = epochs;
      lr_worker.train();
    } else if (*(argv[3]) == '1') {
      std::cout << "start FM " << std::endl;
      xflow::FMWorker fm_worker(argv[1], argv[2]);
      fm_worker.epochs = epochs;
      fm_worker.train();
    } else {
      std::cout << "model " << argv[3] << " is not part of this build (MVM: see DESIGN.md section 8)" << std::endl;
      return 2;
    }
  } catch (const std::exception& e) {
    std::cerr << "xflow_lr: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
