// Small HTTP/1.1 server (own implementation; the reference uses Boost.Beast:
// src/endpoint/FaabricEndpoint.cpp:18-280).  One acceptor + N worker threads,
// keep-alive, Content-Length bodies.
#pragma once

#include <atomic>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace faabric::endpoint {

struct HttpRequest
{
    std::string method;
    std::string target;
    std::map<std::string, std::string> headers;
    std::string body;
};

struct HttpResponse
{
    int status = 200;
    std::string body;
    std::map<std::string, std::string> headers;
};

class HttpRequestHandler
{
  public:
    virtual ~HttpRequestHandler() = default;

    virtual void onRequest(const HttpRequest& request, HttpResponse& response) = 0;
};

enum class EndpointMode
{
    SIGNAL,
    BG_THREAD
};

class FaabricEndpoint
{
  public:
    FaabricEndpoint();

    FaabricEndpoint(int port,
                    int threadCount,
                    std::shared_ptr<HttpRequestHandler> requestHandlerIn);

    FaabricEndpoint(const FaabricEndpoint&) = delete;

    ~FaabricEndpoint();

    // SIGNAL: blocks until SIGINT/SIGTERM; BG_THREAD: returns immediately
    void start(EndpointMode mode = EndpointMode::SIGNAL);

    void stop();

    int getPort() const { return port; }

  private:
    int port;
    int threadCount;
    std::shared_ptr<HttpRequestHandler> requestHandler;
    struct Impl;
    std::unique_ptr<Impl> impl;
};

}
