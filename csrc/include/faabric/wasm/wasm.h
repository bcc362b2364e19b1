// (reference: include/faabric/wasm/wasm.h - a placeholder for the embedder)
#pragma once

int helloFaabricWasm();
