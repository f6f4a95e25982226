// metalchat_b200/facade/mc_kernel_thread.cc — replaces src/kernel_thread.cc of the reference.
//
// Reference: a kernel_queue is one MTL command queue + the current command buffer + a concurrent compute encoder + an event
// chain that orders consecutive buffers (src/kernel_thread.cc:13-56); a kernel_thread is one command buffer of up to
// `capacity` kernels whose std::promise is fulfilled by Metal's completion handler (:123-199).
//
// B200: the queue is the device's ONE in-order CUDA stream (mc_device), a command buffer is an mc_cmdbuf (kernels are
// enqueued on the stream as they are encoded, "commit" records an event, the event chain is the stream order itself).
// Completion handlers run on a worker thread owned by the queue -- the counterpart of Metal's completion thread -- which
// waits for the buffer's event, then releases the kernels' argument tensors and fulfils the promise.  (CUDA host
// callbacks may not call the CUDA API, and releasing a tensor frees device memory.)
#include <condition_variable>
#include <deque>
#include <mutex>
#include <sstream>
#include <thread>

#include <metalchat/kernel_thread.h>

#include "mc_metal_impl.h"


namespace metalchat {


namespace {


/// Completion thread: processes committed command buffers in commit order.
class completion_worker {
public:
    using job_type = std::function<void()>;

    completion_worker()
    : _M_state(std::make_shared<state>())
    {
        _M_thread = std::thread([s = _M_state] { run(*s); });
    }

    ~completion_worker()
    {
        {
            std::scoped_lock lock(_M_state->mu);
            _M_state->stop = true;
        }
        _M_state->cv.notify_all();
        // The last owner may be a tensor released by a completion handler, i.e. this very thread.
        if (std::this_thread::get_id() == _M_thread.get_id()) {
            _M_thread.detach();
        } else {
            _M_thread.join();
        }
    }

    void
    submit(job_type job)
    {
        {
            std::scoped_lock lock(_M_state->mu);
            _M_state->jobs.push_back(std::move(job));
        }
        _M_state->cv.notify_one();
    }

private:
    struct state {
        std::mutex mu;
        std::condition_variable cv;
        std::deque<job_type> jobs;
        bool stop = false;
    };

    static void
    run(state& s)
    {
        for (;;) {
            job_type job;
            {
                std::unique_lock lock(s.mu);
                s.cv.wait(lock, [&] { return s.stop || !s.jobs.empty(); });
                if (s.jobs.empty()) {
                    return; // stop requested and everything committed so far has been completed
                }
                job = std::move(s.jobs.front());
                s.jobs.pop_front();
            }
            job();
        }
    }

    std::shared_ptr<state> _M_state;
    std::thread _M_thread;
};


struct command_buffer {
    mc_cmdbuf* handle = nullptr;
    mc_kernel* pending = nullptr; // kernel of the function being encoded (named at dispatch by the C ABI)
    // A Metal command buffer retains every buffer bound to it until it has completed (setBuffer); temporaries that the encoder
    // copies to the device (kernel_thread.h:127-137) live only through that reference.
    std::vector<metal::shared_buffer> retained;
    std::vector<kernel_callback_type> handlers;
    std::mutex mu;

    ~command_buffer()
    {
        if (handle != nullptr) {
            mc_cmdbuf_release(handle);
        }
    }
};


} // namespace


struct kernel_queue {
    std::size_t id = 0;
    std::size_t capacity = 64;

    metal::shared_device device;
    std::shared_ptr<completion_worker> worker;
    std::shared_ptr<command_buffer> commands;

    kernel_queue() = default;

    kernel_queue(metal::shared_device d, std::size_t thread_capacity)
    : id(0),
      capacity(thread_capacity),
      device(d),
      worker(std::make_shared<completion_worker>()),
      commands(begin(d, thread_capacity))
    {}

    static std::shared_ptr<command_buffer>
    begin(const metal::shared_device& d, std::size_t thread_capacity)
    {
        auto cb = std::make_shared<command_buffer>();
        metal::check(mc_stream_begin(d->handle, thread_capacity, &cb->handle));
        return cb;
    }

    /// The next command buffer of the same queue (src/kernel_thread.cc:35-47); ordering after the previous buffer is the
    /// stream order, no event has to be encoded.
    kernel_queue
    partition() const
    {
        kernel_queue kq = *this;
        kq.id++;
        kq.commands = begin(device, capacity);
        return kq;
    }

    void
    on_completed(kernel_callback_type callback)
    {
        std::scoped_lock lock(commands->mu);
        commands->handlers.push_back(std::move(callback));
    }
};


hardware_function_encoder::hardware_function_encoder(
    std::shared_ptr<kernel_queue> queue_ptr, hardware_function_encoder::allocator_type alloc
)
: _M_allocator(alloc),
  _M_queue(queue_ptr),
  _M_buffer(0),
  _M_name()
{}


// The kernel is remembered until dispatch(): the C ABI binds arguments to slots first and names the kernel when it
// dispatches (mc_dispatch); Metal sets the pipeline state first (src/kernel_thread.cc:69-74).
void
hardware_function_encoder::initialize(const std::string& name, const metal::shared_kernel kernel)
{
    _M_name = name;
    _M_queue->commands->pending = kernel->handle;
}


void
hardware_function_encoder::encode(const void* data, std::size_t size)
{
    metal::check(mc_set_bytes(_M_queue->commands->handle, std::uint32_t(_M_buffer++), data, size));
}


void
hardware_function_encoder::encode(metal::shared_buffer buffer, std::size_t offset)
{
    metal::check(mc_set_buffer(
        _M_queue->commands->handle, std::uint32_t(_M_buffer++), buffer->handle, buffer->offset + offset
    ));
    std::scoped_lock lock(_M_queue->commands->mu);
    _M_queue->commands->retained.push_back(buffer);
}


void
hardware_function_encoder::encode_memory_barrier(metal::shared_buffer buffer)
{
    // needed by the reference because its encoder is concurrent (src/kernel_thread.cc:91-96); a no-op on an in-order stream
    metal::check(mc_barrier(_M_queue->commands->handle, buffer->handle));
}


void
hardware_function_encoder::on_completed(kernel_callback_type callback)
{
    _M_queue->on_completed(callback);
}


void
hardware_function_encoder::dispatch(dim3 grid, dim3 group)
{
    const std::uint32_t g[3] = {std::uint32_t(grid.x), std::uint32_t(grid.y), std::uint32_t(grid.z)};
    const std::uint32_t t[3] = {std::uint32_t(group.x), std::uint32_t(group.y), std::uint32_t(group.z)};
    metal::check(mc_dispatch(_M_queue->commands->handle, _M_queue->commands->pending, g, t), _M_name.c_str());
}


kernel_thread::kernel_thread(const kernel_queue& queue, std::size_t capacity, allocator_type alloc)
: _M_allocator(alloc),
  _M_queue(std::make_shared<kernel_queue>(queue)),
  _M_promise(std::make_shared<promise_type>()),
  _M_future(_M_promise->get_future()),
  _M_size(0),
  _M_capacity(capacity),
  _M_committed(false)
{}


void
kernel_thread::on_completed(kernel_callback_type callback)
{
    _M_queue->on_completed(callback);
}


kernel_thread::~kernel_thread()
{
    // an uncommitted buffer is committed when its thread goes away (src/kernel_thread.cc:154-159)
    try {
        make_ready_at_thread_exit();
    } catch (...) {
    }
}


std::size_t
kernel_thread::size() const
{
    return _M_size;
}


std::size_t
kernel_thread::capacity() const
{
    return _M_capacity;
}


bool
kernel_thread::joinable() const
{
    return (!_M_committed) && (_M_size < _M_capacity);
}


void
kernel_thread::make_ready_at_thread_exit()
{
    if (_M_committed) {
        return;
    }
    _M_committed = true;

    auto commands = _M_queue->commands;
    const mc_status committed = mc_commit(commands->handle);
    const std::string commit_error = committed == MC_OK ? std::string() : std::string(mc_last_error());

    // completion: wait for the buffer's event, run the handlers (they release the kernels' arguments), fulfil the promise
    // with the buffer's error if it has one (src/kernel_thread.cc:134-144)
    _M_queue->worker->submit([commands, promise = _M_promise, committed, commit_error] {
        char err[512] = {0};
        const mc_status status = committed == MC_OK ? mc_wait(commands->handle, err, sizeof(err)) : committed;

        std::vector<kernel_callback_type> handlers;
        std::vector<metal::shared_buffer> retained;
        {
            std::scoped_lock lock(commands->mu);
            handlers.swap(commands->handlers);
            retained.swap(commands->retained);
        }
        for (auto& handler : handlers) {
            handler();
        }
        handlers.clear();
        retained.clear();

        if (status != MC_OK) {
            const std::string what = committed == MC_OK ? std::string(err) : commit_error;
            promise->set_exception(std::make_exception_ptr(std::runtime_error(what)));
        } else {
            promise->set_value();
        }
    });
}


recursive_kernel_thread::recursive_kernel_thread(
    metal::shared_device device, std::size_t thread_capacity
)
: _M_allocator(hardware_memory_allocator(device)),
  _M_queue(std::make_shared<kernel_queue>(device, thread_capacity)),
  _M_thread(std::make_shared<kernel_thread>(*_M_queue, thread_capacity, _M_allocator)),
  _M_thread_capacity(thread_capacity)
{}


std::shared_ptr<kernel_thread>
recursive_kernel_thread::get_this_thread()
{
    if (!_M_thread->joinable()) {
        auto queue = std::make_shared<kernel_queue>(_M_queue->partition());
        auto thread = std::make_shared<kernel_thread>(*queue, _M_thread_capacity, _M_allocator);

        _M_queue.swap(queue);
        _M_thread.swap(thread);
    }
    return _M_thread;
}


} // namespace metalchat
