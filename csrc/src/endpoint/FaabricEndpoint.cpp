// Own HTTP/1.1 server: blocking acceptor + worker pool over a connection
// queue.  Workers keep a connection for its keep-alive lifetime.
#include <faabric/endpoint/FaabricEndpoint.h>
#include <faabric/endpoint/FaabricEndpointHandler.h>
#include <faabric/util/config.h>
#include <faabric/util/logging.h>
#include <faabric/util/queue.h>
#include <faabric/util/string_tools.h>

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <signal.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>

namespace faabric::endpoint {

static const size_t MAX_HEADER_BYTES = 64 * 1024;
static const size_t MAX_BODY_BYTES = (size_t)512 << 20;

struct FaabricEndpoint::Impl
{
    int listenFd = -1;
    std::atomic<bool> running{ false };
    std::thread acceptor;
    std::vector<std::thread> workers;
    faabric::util::Queue<int> connQueue;
    std::mutex liveMx;
    std::vector<int> liveFds;
};

FaabricEndpoint::FaabricEndpoint()
  : FaabricEndpoint(faabric::util::getSystemConfig().endpointPort,
                    faabric::util::getSystemConfig().endpointNumThreads,
                    std::make_shared<FaabricEndpointHandler>())
{}

FaabricEndpoint::FaabricEndpoint(int portIn, int threadCountIn, std::shared_ptr<HttpRequestHandler> handlerIn)
  : port(portIn)
  , threadCount(std::max(1, threadCountIn))
  , requestHandler(std::move(handlerIn))
  , impl(std::make_unique<Impl>())
{}

FaabricEndpoint::~FaabricEndpoint()
{
    stop();
}

static std::string lower(std::string s)
{
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
    return s;
}

static const char* statusText(int code)
{
    switch (code) {
        case 200:
            return "OK";
        case 400:
            return "Bad Request";
        case 404:
            return "Not Found";
        case 405:
            return "Method Not Allowed";
        case 413:
            return "Payload Too Large";
        case 500:
            return "Internal Server Error";
        default:
            return "Unknown";
    }
}

// Returns false when the peer closed / errored
static bool readMore(int fd, std::string& buf, const std::atomic<bool>& running)
{
    char tmp[16384];
    while (running.load(std::memory_order_relaxed)) {
        struct pollfd p = { fd, POLLIN, 0 };
        int pr = ::poll(&p, 1, 200);
        if (pr < 0 && errno == EINTR) {
            continue;
        }
        if (pr < 0) {
            return false;
        }
        if (pr == 0) {
            continue;
        }
        ssize_t n = ::recv(fd, tmp, sizeof(tmp), 0);
        if (n > 0) {
            buf.append(tmp, (size_t)n);
            return true;
        }
        if (n < 0 && (errno == EINTR || errno == EAGAIN)) {
            continue;
        }
        return false;
    }
    return false;
}

static bool writeAll(int fd, const std::string& data)
{
    size_t off = 0;
    while (off < data.size()) {
        ssize_t n = ::send(fd, data.data() + off, data.size() - off, MSG_NOSIGNAL);
        if (n < 0) {
            if (errno == EINTR) {
                continue;
            }
            return false;
        }
        off += (size_t)n;
    }
    return true;
}

static void serveConnection(int fd, HttpRequestHandler& handler, const std::atomic<bool>& running)
{
    std::string buf;
    for (;;) {
        // ---- headers ----
        size_t headerEnd;
        while ((headerEnd = buf.find("\r\n\r\n")) == std::string::npos) {
            if (buf.size() > MAX_HEADER_BYTES || !readMore(fd, buf, running)) {
                return;
            }
        }
        HttpRequest req;
        {
            size_t lineEnd = buf.find("\r\n");
            std::string start = buf.substr(0, lineEnd);
            size_t s1 = start.find(' ');
            size_t s2 = start.rfind(' ');
            if (s1 == std::string::npos || s2 == s1) {
                writeAll(fd, "HTTP/1.1 400 Bad Request\r\nContent-Length: 0\r\nConnection: close\r\n\r\n");
                return;
            }
            req.method = start.substr(0, s1);
            req.target = start.substr(s1 + 1, s2 - s1 - 1);
            size_t pos = lineEnd + 2;
            while (pos < headerEnd) {
                size_t e = buf.find("\r\n", pos);
                std::string line = buf.substr(pos, e - pos);
                size_t colon = line.find(':');
                if (colon != std::string::npos) {
                    std::string k = lower(line.substr(0, colon));
                    size_t vs = line.find_first_not_of(" \t", colon + 1);
                    req.headers[k] = vs == std::string::npos ? "" : line.substr(vs);
                }
                pos = e + 2;
            }
        }
        size_t bodyStart = headerEnd + 4;
        size_t contentLength = 0;
        if (auto it = req.headers.find("content-length"); it != req.headers.end()) {
            contentLength = (size_t)std::strtoull(it->second.c_str(), nullptr, 10);
        }
        if (contentLength > MAX_BODY_BYTES) {
            writeAll(fd, "HTTP/1.1 413 Payload Too Large\r\nContent-Length: 0\r\nConnection: close\r\n\r\n");
            return;
        }
        if (auto it = req.headers.find("expect"); it != req.headers.end() && lower(it->second) == "100-continue") {
            writeAll(fd, "HTTP/1.1 100 Continue\r\n\r\n");
        }
        while (buf.size() < bodyStart + contentLength) {
            if (!readMore(fd, buf, running)) {
                return;
            }
        }
        req.body = buf.substr(bodyStart, contentLength);
        buf.erase(0, bodyStart + contentLength);

        bool keepAlive = true;
        if (auto it = req.headers.find("connection"); it != req.headers.end()) {
            keepAlive = lower(it->second) != "close";
        }

        HttpResponse resp;
        resp.headers["Content-Type"] = "text/plain";
        try {
            if (req.method == "OPTIONS") {
                resp.status = 200;
                resp.headers["Access-Control-Allow-Origin"] = "*";
                resp.headers["Access-Control-Allow-Methods"] = "GET,POST,PUT,OPTIONS";
                resp.headers["Access-Control-Allow-Headers"] = "User-Agent,Content-Type";
            } else {
                handler.onRequest(req, resp);
            }
        } catch (std::exception& e) {
            SPDLOG_ERROR("HTTP handler threw: {}", e.what());
            resp.status = 500;
            resp.body = std::string("Caught exception: ") + e.what();
        }

        std::string out = "HTTP/1.1 " + std::to_string(resp.status) + " " + statusText(resp.status) + "\r\n";
        for (const auto& [k, v] : resp.headers) {
            out += k + ": " + v + "\r\n";
        }
        out += "Content-Length: " + std::to_string(resp.body.size()) + "\r\n";
        out += keepAlive ? "Connection: keep-alive\r\n\r\n" : "Connection: close\r\n\r\n";
        out += resp.body;
        if (!writeAll(fd, out) || !keepAlive) {
            return;
        }
    }
}

static std::atomic<bool> gotSignal{ false };

static void onSignal(int)
{
    gotSignal.store(true);
}

void FaabricEndpoint::start(EndpointMode mode)
{
    if (impl->running.exchange(true)) {
        return;
    }
    int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) {
        throw std::runtime_error("Endpoint socket() failed");
    }
    int one = 1;
    ::setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = htonl(INADDR_ANY);
    addr.sin_port = htons((uint16_t)port.load());
    if (::bind(fd, (sockaddr*)&addr, sizeof(addr)) != 0 || ::listen(fd, 1024) != 0) {
        ::close(fd);
        impl->running = false;
        throw std::runtime_error("Endpoint could not bind to port " + std::to_string(port.load()) + ": " + strerror(errno));
    }
    if (port == 0) {
        socklen_t len = sizeof(addr);
        ::getsockname(fd, (sockaddr*)&addr, &len);
        port = ntohs(addr.sin_port);
    }
    impl->listenFd = fd;
    SPDLOG_INFO("Starting HTTP endpoint on {}, {} threads", port.load(), threadCount);

    Impl* im = impl.get();
    auto handler = requestHandler;
    for (int i = 0; i < threadCount; i++) {
        im->workers.emplace_back([im, handler] {
            for (;;) {
                int cfd;
                try {
                    cfd = im->connQueue.dequeue(500);
                } catch (faabric::util::QueueTimeoutException&) {
                    continue;
                }
                if (cfd < 0) {
                    return;
                }
                {
                    std::lock_guard<std::mutex> lk(im->liveMx);
                    im->liveFds.push_back(cfd);
                }
                serveConnection(cfd, *handler, im->running);
                {
                    std::lock_guard<std::mutex> lk(im->liveMx);
                    im->liveFds.erase(std::remove(im->liveFds.begin(), im->liveFds.end(), cfd), im->liveFds.end());
                }
                ::close(cfd);
            }
        });
    }
    im->acceptor = std::thread([im] {
        while (im->running.load()) {
            struct pollfd p = { im->listenFd, POLLIN, 0 };
            int pr = ::poll(&p, 1, 200);
            if (pr <= 0) {
                continue;
            }
            int cfd = ::accept4(im->listenFd, nullptr, nullptr, SOCK_CLOEXEC);
            if (cfd < 0) {
                continue;
            }
            int on = 1;
            ::setsockopt(cfd, IPPROTO_TCP, TCP_NODELAY, &on, sizeof(on));
            im->connQueue.enqueue(cfd);
        }
    });

    if (mode == EndpointMode::SIGNAL) {
        gotSignal = false;
        struct sigaction sa{};
        sa.sa_handler = onSignal;
        sigaction(SIGINT, &sa, nullptr);
        sigaction(SIGTERM, &sa, nullptr);
        sigaction(SIGQUIT, &sa, nullptr);
        while (!gotSignal.load() && impl->running.load()) {
            ::usleep(100 * 1000);
        }
        SPDLOG_INFO("Received signal, shutting down HTTP endpoint");
        stop();
    }
}

void FaabricEndpoint::stop()
{
    if (!impl || !impl->running.exchange(false)) {
        return;
    }
    SPDLOG_DEBUG("Shutting down endpoint on {}", port.load());
    if (impl->acceptor.joinable()) {
        impl->acceptor.join();
    }
    for (size_t i = 0; i < impl->workers.size(); i++) {
        impl->connQueue.enqueue(-1);
    }
    {
        // Kick workers blocked on idle keep-alive connections
        std::lock_guard<std::mutex> lk(impl->liveMx);
        for (int fd : impl->liveFds) {
            ::shutdown(fd, SHUT_RDWR);
        }
    }
    for (auto& w : impl->workers) {
        if (w.joinable()) {
            w.join();
        }
    }
    impl->workers.clear();
    // Drain connections nobody picked up
    for (;;) {
        try {
            int fd = impl->connQueue.dequeue(1);
            if (fd >= 0) {
                ::close(fd);
            }
        } catch (...) {
            break;
        }
    }
    ::close(impl->listenFd);
    impl->listenFd = -1;
}

void FaabricEndpointHandler::onRequest(const HttpRequest& request, HttpResponse& response)
{
    SPDLOG_ERROR("Worker HTTP handler received a request ({} {}), this is not supported", request.method, request.target);
    response.status = 400;
    response.headers["Server"] = "Worker endpoint";
    response.body = "Worker HTTP handler does not accept requests; talk to the planner";
}

}
